#!/bin/bash
# build the C-ABI library; non-zero exit (and no GPU time spent) when nvcc fails
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py > /tmp/mmsa_build.log 2>&1 || { grep -E "error|Error" -A4 /tmp/mmsa_build.log | head -40; exit 1; }
tail -1 /tmp/mmsa_build.log
