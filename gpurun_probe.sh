#!/bin/bash
GCKPP_PROFILE=1 timeout 300 python tools/smem_one.py 444 2>&1 | tail -4
