#!/bin/bash
# sweep of the polling parameters of the dataflow sweeps (B200_FLOW_NEAR, B200_FLOW_PER_SIGNAL, B200_FLOW_SLEEP)
for near in 16 64 256; do for per in 2 4 8 16; do
  a=$(B200_FLOW_NEAR=$near B200_FLOW_PER_SIGNAL=$per timeout 100 python profiles/prof_driver.py 1 2>&1 | tail -1)
  b=$(B200_FLOW_NEAR=$near B200_FLOW_PER_SIGNAL=$per timeout 100 python profiles/prof_driver.py 2 2>&1 | tail -1)
  echo "near=$near per=$per | $a | $b"
done; done
