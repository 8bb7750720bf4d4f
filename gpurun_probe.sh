#!/bin/bash
( GCKPP_PROFILE=1 timeout 120 python tools/smem_one.py 444 ) 2>&1 | tail -3
