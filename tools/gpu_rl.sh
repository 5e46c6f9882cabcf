#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/rainbow_loop_profile.py > gpurun_out/rainbow_loop.log 2>&1; tail -5 gpurun_out/rainbow_loop.log
