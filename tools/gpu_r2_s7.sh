#!/bin/bash
# Round 2, session 7: whole GPU suite after the multi-rank / legacy / audit work, smoke.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_s7.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r02_pytest_s7.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
