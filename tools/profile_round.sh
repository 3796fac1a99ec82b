#!/bin/bash
# Round profile recipe (run on the GPU box through gpurun): bench lines, ncu launch list of the bench command, full
# captures of the dominant kernels.  Outputs under gpurun_out/; tools/summarize_ncu.py turns them into profiles/*.txt.
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/launches_${TAG}.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_snorm_batch -s 3 -c 1 -f -o gpurun_out/prof_snorm_${TAG} \
    python bench.py --cases 148 --contact-cases 148 --steps 1 --warmup 3 --skip-extra --cpu-seconds 0 > gpurun_out/prof_snorm_${TAG}.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_contac_batch -s 3 -c 1 -f -o gpurun_out/prof_contac_${TAG} \
    python bench.py --cases 148 --contact-cases 148 --steps 1 --warmup 3 --skip-extra --cpu-seconds 0 > gpurun_out/prof_contac_${TAG}.out 2>&1
ncu --set full --clock-control none -k regex:k_lg_ -s 12 -c 3 -f -o gpurun_out/prof_large_${TAG} \
    python tools/large_product_only.py > gpurun_out/prof_large_${TAG}.out 2>&1
ncu --set full --clock-control none -k regex:k_lg_contac -c 1 -f -o gpurun_out/prof_gd_${TAG} \
    python tools/gd_timing.py 2c > gpurun_out/prof_gd_${TAG}.out 2>&1
for f in prof_snorm prof_contac prof_large prof_gd; do
    ncu -i gpurun_out/${f}_${TAG}.ncu-rep --page raw --csv > gpurun_out/${f}_${TAG}.raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_snorm_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_snorm_${TAG}.source.csv 2>/dev/null
ncu -i gpurun_out/prof_contac_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_contac_${TAG}.source.csv 2>/dev/null
# source counters of the SteadyGS sweeps (one wave of hertz-91 cases): input of tools/chain_profile.py -> profiles/steadygs_chain_*.txt
ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:k_contac_batch -c 1 -f \
    -o gpurun_out/prof_gs_${TAG} python tools/steady_timing_h91.py 148 > gpurun_out/prof_gs_${TAG}.out 2>&1
ncu -i gpurun_out/prof_gs_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_gs_${TAG}.source.csv 2>/dev/null
gzip -f gpurun_out/prof_gs_${TAG}.source.csv
# race / memory checks of the sweeps
compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_steady.py > gpurun_out/sanitize_racecheck_${TAG}.log 2>&1
compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_steady.py > gpurun_out/sanitize_memcheck_${TAG}.log 2>&1
ls -la gpurun_out/*.ncu-rep
# keep the merge under the 64 MiB limit
for f in gpurun_out/*.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 20000000 ]; then rm -f $f; fi; done
gzip -f gpurun_out/prof_snorm_${TAG}.source.csv gpurun_out/prof_contac_${TAG}.source.csv
du -sh gpurun_out
