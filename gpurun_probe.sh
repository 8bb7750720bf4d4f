#!/bin/bash
( GCKPP_PROFILE=1 timeout 120 python tools/smem_one.py 296 ) 2>&1 | tail -3
( timeout 300 python tools/smem_debug.py small ) 2>&1 | tail -3
