#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag_side.log
: > $L
for kind in unet nested; do
  MDM_SIDE_WGRAD=0 MDM_NO_GRAPH=1 python tests/diag_side.py $kind >> $L 2>&1
  echo "== $kind eager side=1 vs side=0" >> $L
  MDM_SIDE_WGRAD=1 MDM_NO_GRAPH=1 python tests/diag_side.py $kind /tmp/diag_${kind}_0_1.pt >> $L 2>&1
  echo "== $kind graph side=1 vs eager side=0" >> $L
  MDM_SIDE_WGRAD=1 python tests/diag_side.py $kind /tmp/diag_${kind}_0_1.pt >> $L 2>&1
  echo "== $kind graph side=0 vs eager side=0" >> $L
  MDM_SIDE_WGRAD=0 python tests/diag_side.py $kind /tmp/diag_${kind}_0_1.pt >> $L 2>&1
done
cat $L
