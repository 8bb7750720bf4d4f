#!/bin/bash
for f in libvar_chain2 libvar_chain0; do
  GCKPP_B200_LIB=$PWD/geos_chem_b200/$f.so timeout 200 python tools/variant_bench.py own 2>&1 | tail -1
done
