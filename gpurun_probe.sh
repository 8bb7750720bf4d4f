#!/bin/bash
timeout 120 python tools/hg_debug.py 2>&1 | tail -6
