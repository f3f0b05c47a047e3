#!/bin/bash
# Same-box A/B of the inverse acquisition kernel's launch choices (profiles/acq_r02_split_ab.log):
#   GPSB200_ACQ_QUAD=0  4-CTA form only      =1  quad form only      =2  quad for the whole waves + 4-CTA for the rest
#   (unset)             the launcher's own choice (gr_acq_run_dev)
# configs[1] at 512 / 128 / 16 / 4 recordings, configs[3] on 16 recordings: all bins, a quarter shard, a one-eighth shard.
# usage (on a B200): bash tools/ab_inverse_forms.sh > gpurun_out/split_ab.log 2>&1
for n in 512 128 16 4; do for q in 0 1 2 auto; do
  echo "quad=$q n=$n"
  if [ $q = auto ]; then python tools/acq_time.py $n 10; else GPSB200_ACQ_QUAD=$q python tools/acq_time.py $n 10; fi
done; done
for sh in 1 4 8; do for q in 0 1 2 auto; do
  echo "fine shards=$sh quad=$q"
  if [ $q = auto ]; then python tools/prof_fine.py 16 5 $sh; else GPSB200_ACQ_QUAD=$q python tools/prof_fine.py 16 5 $sh; fi
done; done
