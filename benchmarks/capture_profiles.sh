#!/bin/bash
# ncu evidence for profiles/: run ON the GPU box (gpurun). Full captures are summarised there and deleted -- only text
# comes back (gpurun_out/ is capped at 64 MiB).   usage: bash benchmarks/capture_profiles.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
MEM='Kernel Name|gpu__time_duration.sum|launch__registers_per_thread$|l1tex__t_(sectors|requests)_pipe_lsu_mem_global_op_(ld|red|st).sum$|l1tex__t_sector_hit_rate|lts__t_requests.sum$|lts__t_sectors_srcunit_tex_op_(read|red|write).sum$|lts__t_sector_hit_rate.pct|dram__bytes_(read|write).sum$|l1tex__m_l1tex2xbar_req_cycles_active.avg.pct|lts__t_tag_requests|smsp__issue_active.avg.pct|sm__warps_active.avg.pct'
summ() {  # rep name
  python benchmarks/ncu_summary.py $OUT/$1.ncu-rep > $OUT/${TAG}_ncu_$1_summary.txt 2>&1
  echo "---- memory system (requests, sectors, hit rates, DRAM bytes per launch) ----" >> $OUT/${TAG}_ncu_$1_summary.txt
  python benchmarks/ncu_dump.py $OUT/$1.ncu-rep "$MEM" >> $OUT/${TAG}_ncu_$1_summary.txt 2>&1
  ncu -i $OUT/$1.ncu-rep --page details --print-details all 2>/dev/null | grep -E "^  void|Memory Throughput Breakdown|L1: M L1tex2xbar|L2: T Tag|L1: Data Pipe Lsu|L2: D Atomic|L1: Lsu Writeback|DRAM: Cycles" >> $OUT/${TAG}_ncu_$1_summary.txt
  rm -f $OUT/$1.ncu-rep
}
$NCU -k regex:"latent_bwd_tiled|latent_fwd_tiled" -s 2 -c 2 -f -o $OUT/tiled2d python benchmarks/ncu_2d.py --reps 2 > $OUT/ncu.log 2>&1; summ tiled2d
$NCU -k regex:"entropy_kernel|sga_" -s 1 -c 1 -f -o $OUT/entropy python benchmarks/ncu_2d.py --reps 2 --entropy >> $OUT/ncu.log 2>&1; summ entropy
$NCU -k regex:"latent_fwd3d|latent_bwd3d|latent_bwd_tiled" -s 3 -c 3 -f -o $OUT/nerf3d python benchmarks/ncu_3d.py --reps 2 --sorted 128 >> $OUT/ncu.log 2>&1; summ nerf3d
$NCU -k regex:"plan_" -s 3 -c 3 -f -o $OUT/plan3d python benchmarks/ncu_3d.py --reps 2 --sorted 128 --what fwd >> $OUT/ncu.log 2>&1; summ plan3d
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > $OUT/${TAG}_launches_bench.log 2>&1
ls -la $OUT
