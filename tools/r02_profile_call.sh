#!/bin/bash
# One gpurun call (1 GPU): the round-2 bench lines and ncu evidence kept under profiles/.  Run from the repo root:
#   gpurun --timeout 2400 -- 'bash tools/r02_profile_call.sh'
mkdir -p gpurun_out
O=gpurun_out
# 1. the bench lines (never under a profiler)
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1_final.json 2> $O/bench_final.err; tail -c 600 $O/r02_bench_n1_final.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2>> $O/bench_final.err; tail -c 400 $O/r02_bench_reference_arm.json; echo
timeout 300 python tools/band_bench.py --loci 5000 > $O/r02_band_bench.json 2> $O/band_bench.err; tail -c 700 $O/r02_band_bench.json; echo
# 2. launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_ncu_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
# 3. the folded kernel, full set, one launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vtx_k_sw_fold -s 3 -c 1 -f -o $O/r02_fold_final python bench.py --loci 20000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_fold.log 2>&1
python tools/ncu_kernel_metrics.py $O/r02_fold_final.ncu-rep --kernel vtx_k_sw_fold --what "ncu --set full --clock-control none, bench.py --loci 20000 --steps 1 --warmup 1 (final round-2 build)" > $O/r02_fold_final_metrics.json 2>> $O/ncu_fold.log
tail -c 300 $O/ncu_fold.log; echo
# 4. the staging kernels (device inflate, record walk, parse, per-locus fetch + filters, tag extraction) inside the CLI
D=/tmp/vtx_prof_ds; mkdir -p $D
python -c "
import sys; sys.path.insert(0, '.')
from vartrix_b200 import synth_files
synth_files.write_dataset_fast('$D', n_loci=20000, n_barcodes=5000, depth=50, read_len=150, seed=2)
" 2> $O/ds.err
timeout 900 ncu --set full --clock-control none -k 'regex:bgzf|walk|parse|locus_cands|read_emit|widen' -c 60 -f -o $O/r02_stage_kernels vartrix_b200/bin/vartrix_b200 -v $D/variants.vcf -b $D/reads.bam -f $D/genome.fa -c $D/barcodes.tsv -o $D/out.mtx --threads 4 --gpu-stage --shard-loci 2500 > $O/ncu_stage.log 2>&1
python tools/ncu_kernel_metrics.py $O/r02_stage_kernels.ncu-rep --what "ncu --set full, vartrix_b200 --gpu-stage --shard-loci 2500 on a 1 M-read synthetic BAM (first 60 staging launches)" > $O/r02_stage_kernels_metrics.json 2>> $O/ncu_stage.log
rm -f $O/r02_stage_kernels.ncu-rep        # the metrics JSON is what is kept
tail -c 300 $O/ncu_stage.log; echo
ls -la $O | tail -20
