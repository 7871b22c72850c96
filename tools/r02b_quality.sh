#!/bin/bash
# quality sweep on the shipping routes (12 committed runs x 3 seeds)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out


timeout 900 python tools/quality_sweep.py > gpurun_out/r02b_quality_sweep.json 2> gpurun_out/r02b_quality_sweep.log
echo "quality rc=$?"; tail -16 gpurun_out/r02b_quality_sweep.log
