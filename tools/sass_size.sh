#!/bin/bash
# SASS instruction count per kernel of the built library
cuobjdump -sass "${1:-clfacedetection_b200/libclfd_b200.so}" 2>/dev/null | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {n[name]++} END {for (k in n) print n[k], k}' | sort -n
