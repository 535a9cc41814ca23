#!/bin/bash
# Registers / stack of every kernel in the built library (cuobjdump is the authority: the
# `ptxas -v` lines are easy to pair with the wrong entry when device functions are not inlined).
cuobjdump -res-usage "${1:-clfacedetection_b200/libclfd_b200.so}" 2>/dev/null | grep -A1 "^ Function" | grep -v "^--" | paste - - |
  sed 's/^ Function //; s/SHARED.*//' | c++filt -_ 2>/dev/null | awk '{print}' | cut -c1-200
