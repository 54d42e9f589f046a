#!/bin/bash
# tools/gpu_job.sh TAG -- the standard measurement job of one gpurun call: GPU tests, bench, ncu captures (exported as CSV)
tag=${1:-r02x}
o=gpurun_out
python -c "import bench; print(bench.source_sha())" > $o/sha_$tag.txt
python -m pytest tests -m gpu -q 2>&1 | tail -150 > $o/${tag}_pytest.txt; tail -3 $o/${tag}_pytest.txt
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; head -c 200 $o/${tag}_bench.json; echo
B="python bench.py --n-env 2048 --steps 2 --warmup 1 --no-path-b --no-cpu-baseline --no-compact --parity-scans 0"
$B > $o/${tag}_plainA.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qd_scan_fast -s 3 -c 1 -o $o/prof_${tag}_fast8 $B > $o/${tag}_ncuA.log 2>&1
T="python tools/tunnel_time.py --cases 8:64 --steps 1"
$T > $o/${tag}_plainB.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/${tag}_launches_tunnel.csv $T > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qd_tunnel_select2 -s 2 -c 1 -o $o/prof_${tag}_select8 $T > $o/${tag}_ncuS.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qd_tunnel_eigen2 -s 2 -c 1 -o $o/prof_${tag}_eigen8 $T > $o/${tag}_ncuE.log 2>&1
T4="python tools/tunnel_time.py --cases 4:256 --steps 1"
$T4 > $o/${tag}_plainB4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qd_tunnel_select2 -s 2 -c 1 -o $o/prof_${tag}_select4 $T4 > $o/${tag}_ncuS4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qd_tunnel_eigen2 -s 2 -c 1 -o $o/prof_${tag}_eigen4 $T4 > $o/${tag}_ncuE4.log 2>&1
QDSIM_EIGEN=householder ncu --set full --clock-control none --import-source on -k regex:qd_tunnel_eigen_kernel -s 2 -c 1 -o $o/prof_${tag}_hh8 $T > $o/${tag}_ncuH.log 2>&1
QDSIM_SELECT=block ncu --set full --clock-control none --import-source on -k regex:qd_tunnel_select_kernel -s 2 -c 1 -o $o/prof_${tag}_blk8 $T > $o/${tag}_ncuK.log 2>&1
tools/ncu_export.sh $o/prof_${tag}_blk8 $o/prof_${tag}_hh8 $o/prof_${tag}_fast8 $o/prof_${tag}_select8 $o/prof_${tag}_eigen8 $o/prof_${tag}_select4 $o/prof_${tag}_eigen4
python tools/sweep.py > $o/${tag}_sweep.json 2> $o/${tag}_sweep.err
QDSIM_TRACE=1 python tools/latency.py > $o/${tag}_latency.json 2> $o/${tag}_latency.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2>/dev/null
python bench.py --path B --n-dot 4 --n-env 1024 --no-compact --parity-scans 0 > $o/${tag}_bench_pathB_4dot.json 2> $o/${tag}_bench_pathB.err
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total,power.limit --format=csv > $o/${tag}_gpu_info.csv
python tools/shell_breakdown.py > $o/${tag}_shell_breakdown.txt 2>&1
du -sh $o
