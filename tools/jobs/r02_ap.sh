#!/bin/bash
# round 2, call AP: smoke() on the final tree (the last seconds of the GPU budget)
timeout 28 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v libpng | tail -3
