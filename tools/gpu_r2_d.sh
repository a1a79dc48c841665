#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python tools/persist_timeline.py > gpurun_out/d_timeline.log 2>&1
grep "====\|globaltimer\|lifetime\|dW launch" gpurun_out/d_timeline.log
