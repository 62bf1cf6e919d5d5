#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -150 ) > gpurun_out/r2_pytest2.log 2>&1
tail -120 gpurun_out/r2_pytest2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
