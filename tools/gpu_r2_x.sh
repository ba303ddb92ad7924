#!/bin/bash
timeout 600 python tools/sweep.py --config c4 --variants "4,2,2;4,3,2;4,1,2;4,2,3;4,3,3;4,2,1;4,3,1" --reps 10 2>&1 | tail -9
timeout 600 python tools/sweep.py --config c3 --variants "4,2,2;4,3,2;4,2,3;4,3,3" --reps 10 2>&1 | tail -5
