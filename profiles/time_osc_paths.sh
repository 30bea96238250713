#!/bin/bash
# config 3 on the three forms of the sustain regime (not a bench number): tcgen05 + mma.sync fill (default), mma.sync only, FFMA2
python profiles/time_osc.py "default (tcgen05 for the steady 16,384-sample chunks)"
CA_OSC_PATH=mma python profiles/time_osc.py "CA_OSC_PATH=mma (mma.sync m16n8k16 only)"
CA_OSC_PATH=fp32 python profiles/time_osc.py "CA_OSC_PATH=fp32 (FFMA2)"
