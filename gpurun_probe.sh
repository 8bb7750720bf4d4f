#!/bin/bash
timeout 200 python tools/variant_bench.py own 2>&1 | tail -2
