#!/bin/bash
# A/B of the working-tree library against scratch/exp/lib_prev.so: bitwise outputs + cold / warm step time
python scratch/r3_bits.py dump /tmp/new.pt | tail -1; DGPMP2_LIB=scratch/exp/lib_prev.so python scratch/r3_bits.py dump /tmp/prev.pt | tail -1
python scratch/r3_bits.py cmp /tmp/prev.pt /tmp/new.pt
echo new; python scratch/r3_cold.py 64 1024; echo prev; DGPMP2_LIB=scratch/exp/lib_prev.so python scratch/r3_cold.py 64 1024
