#!/bin/bash
# ad-hoc GPU probe used during development
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
