#!/bin/bash
# forward-kernel variants: parity first, then device time (run on a B200 box)
set -u
echo "== parity, staged kernels on"
NTTB200_CT_H=1 NTTB200_POLY_CT_STAGED=1 python -m pytest tests -x -q -m gpu -k "ct or polymul or round or rns" 2>&1 | tail -2
NTTB200_CT_H=2 NTTB200_POLY_CT_STAGED=1 python -m pytest tests -x -q -m gpu -k "ct_vs" 2>&1 | tail -1
echo "== N=4096: CT_H 0 / 1 / 2, CT_LD 2"
for m in 0 1 2; do NTTB200_CT_H=$m python tools/dev/gs_time.py 12 ct 28; done
NTTB200_CT_LD=2 python tools/dev/gs_time.py 12 ct 28
for m in 0 1 2; do NTTB200_CT_H=$m python tools/dev/gs_time.py 12 ct 26; done
echo "== N=2^13..2^15: first version / staged"
python tools/dev/gs_time.py 13,14,15 ct 28
NTTB200_POLY_CT_STAGED=1 python tools/dev/gs_time.py 13,14,15 ct 28
NTTB200_POLY_CT_STAGED=1 NTTB200_NO_L4=1 python tools/dev/gs_time.py 13,14,15 ct 28
echo "== N=512..2048 forward: classic / 4q-lazy"
python tools/dev/gs_time.py 9,10,11 ct 26
NTTB200_SMALL_CT_L4=1 python -m pytest tests -x -q -m gpu -k "small_n_ct or ct_vs" 2>&1 | tail -1
NTTB200_SMALL_CT_L4=1 python tools/dev/gs_time.py 9,10,11 ct 26
