#!/bin/bash
# One GPU-box call that produces everything profiles/ keeps for a round: the plain bench line, the launch list of the
# same bench command, and one ncu --set full capture per hot kernel (each ncu run directly after the same command has
# exited 0 without ncu).  Usage (under gpurun): bash tools/profile_round.sh r02
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err || { echo "bench failed"; tail -5 $O/${R}_bench_n1.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/plain_bench_short.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_bench.log 2>&1
python tools/fast_loop.py 2000000 3 > $O/plain_fast.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:validity_fast -s 6 -c 2 -o $O/${R}_validity_fast \
      python tools/fast_loop.py 2000000 3 > $O/ncu_fast.log 2>&1
python tools/edt_once.py 3 > $O/plain_edt.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:edt_ -s 8 -c 8 -o $O/${R}_edt \
      python tools/edt_once.py 3 > $O/ncu_edt.log 2>&1
B200DF_EDT_SLABS=1 python tools/edt_once.py 3 > $O/plain_edt1.log 2>&1 &&
  B200DF_EDT_SLABS=1 ncu --set full --clock-control none --import-source on -k regex:edt_ -s 2 -c 2 -o $O/${R}_edt_one_slab \
      python tools/edt_once.py 3 > $O/ncu_edt1.log 2>&1
python tools/grad_loop.py > $O/plain_grad.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gradients_gather -s 3 -c 1 -o $O/${R}_gradients \
      python tools/grad_loop.py > $O/ncu_grad.log 2>&1
# summaries are made here, on the box: only 64 MiB of gpurun_out/ travel back, and the validity report alone is half that
python profiles/summarize.py rep $O/${R}_validity_fast.ncu-rep $O/${R}_validity_fast_full.txt
python profiles/summarize.py traffic $O/${R}_validity_fast.ncu-rep $O/${R}_validity_traffic.json 2000000
python profiles/summarize.py rep $O/${R}_edt.ncu-rep $O/${R}_edt_full.txt
python profiles/summarize.py rep $O/${R}_edt_one_slab.ncu-rep $O/${R}_edt_one_slab_full.txt
python profiles/summarize.py rep $O/${R}_gradients.ncu-rep $O/${R}_gradients_full.txt
python profiles/summarize.py list $O/${R}_launches_bench.csv $O/${R}_launches_bench.txt
ncu -i $O/${R}_validity_fast.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${R}_validity_fast_source.csv.gz
rm -f $O/${R}_validity_fast.ncu-rep $O/${R}_edt.ncu-rep
for f in $O/plain_fast.log $O/plain_edt.log $O/plain_edt1.log $O/plain_grad.log; do tail -n 1 $f; done
du -sh $O; ls -la $O | tail -24
