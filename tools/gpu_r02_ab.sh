#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02ab}
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
