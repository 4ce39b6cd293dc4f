#!/bin/bash
cd tools/ubench/lto_probe
for fb in cb_100a.fatbin cb_100.fatbin; do
  for mode in 1 2 3; do
    echo "== $fb mode $mode"
    ./probe $fb pcsv_load pcsv_store 4096 8 $mode 2>&1 | tail -4
  done
done
echo "== big"
./probe cb_100.fatbin pcsv_load pcsv_store 32768 512 3 2>&1 | tail -3
./probe cb_100a.fatbin pcsv_load pcsv_store 262144 512 3 2>&1 | tail -3
