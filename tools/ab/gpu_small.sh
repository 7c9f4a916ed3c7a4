#!/bin/bash
SARLACC_SPECULATE=1 python tools/ab/gpu_small.py
SARLACC_SPECULATE=0 python tools/ab/gpu_small.py
