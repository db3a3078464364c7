#!/bin/bash
# One gpurun call = one fresh box: run everything that needs a GPU in one go and bring the logs back in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh tests ncu libbar attn'
set -u
mkdir -p gpurun_out
export PYTHONPATH="$PWD:${PYTHONPATH:-}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
md5sum diffusers_b200/_C/libb200diff.so | tee -a gpurun_out/nvsmi.txt
for step in "$@"; do
  t0=$(date +%s)
  case "$step" in
    tests)   timeout 1100 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ;;
    tests_noflux) timeout 900 python -m pytest tests -m gpu -x -q -s --deselect tests/test_full_size_parity_gpu.py::test_flux_full_size_parity > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ;;
    newtests) timeout 900 python -m pytest tests/test_full_size_parity_gpu.py tests/test_dropin_reference_gpu.py -m gpu -q -s > gpurun_out/pytest_new.log 2>&1; echo "new tests rc=$?" ;;
    ncu)     timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_targets python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1; echo "ncu rc=$?" ;;
    libbar)  timeout 600 python tools/library_bar.py --json gpurun_out/library_bar.json > gpurun_out/library_bar.txt 2>&1; echo "libbar rc=$?" ;;
    libquick) timeout 400 python tools/library_bar.py --only gemm,lnfold,attention --json gpurun_out/library_bar_quick.json > gpurun_out/library_bar_quick.txt 2>&1; echo "libquick rc=$?" ;;
    attn)    timeout 300 python tools/bench_attention.py > gpurun_out/bench_attention.txt 2>&1; echo "attn rc=$?" ;;
    attnpoly) for p in 0 1 3; do echo "== B200_ATTN_POLY=$p"; B200_ATTN_POLY=$p timeout 200 python tools/bench_attention.py d64; done > gpurun_out/bench_attention_poly.txt 2>&1; echo "attnpoly rc=$?" ;;
    attnpoly128) for p in 1 2 3; do echo "== B200_ATTN_POLY128=$p"; B200_ATTN_POLY128=$p timeout 200 python tools/bench_attention.py d128; done > gpurun_out/bench_attention_poly128.txt 2>&1; echo "attnpoly128 rc=$?" ;;
    timeline) for m in pf hot; do MODE=$m timeout 120 python tools/gemm_timeline.py 2048 1280 1280; done > gpurun_out/gemm_timeline.txt 2>&1; echo "timeline rc=$?" ;;
    benchsdxl) timeout 600 python bench.py --workload sdxl --no-cpu-baseline --no-reference-cuda > gpurun_out/bench_sdxl.json 2> gpurun_out/bench_sdxl.err; echo "bench sdxl rc=$?" ;;
    attncheck) timeout 400 python tools/diag_ops.py --inproc $(python -c "import sys; sys.path.insert(0,'.'); from tools import diag_ops; print(' '.join(c for c in diag_ops.CASES if c.startswith('attn_')))") > gpurun_out/attn_check.txt 2>&1; echo "attncheck rc=$?"; tail -3 gpurun_out/attn_check.txt ;;
    ncuattn) timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attention -f -o gpurun_out/r2_attn python tools/ncu_targets.py > gpurun_out/ncu_attn.log 2>&1; echo "ncuattn rc=$?" ;;
    gemm)    timeout 600 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; echo "gemm rc=$?" ;;
    bench)   timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" ;;
    benchref) timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" ;;
    flux)    timeout 900 python bench.py --workload flux --steps 5 > gpurun_out/bench_flux.json 2> gpurun_out/bench_flux.err; echo "flux rc=$?" ;;
    vae)     timeout 600 python bench.py --workload vae > gpurun_out/bench_vae.json 2> gpurun_out/bench_vae.err; echo "vae rc=$?" ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_forward.py ncu > gpurun_out/launches.log 2>&1; echo "launches rc=$?" ;;
    ncuhot_*) t=${step#ncuhot_}; timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section LaunchStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --import-source on --clock-control none --profile-from-start off --warp-sampling-interval 0 -c 3 -f -o gpurun_out/r2_hot_$t python tools/ncu_hot.py $t > gpurun_out/ncu_hot_$t.log 2>&1; echo "ncuhot $t rc=$?" ;;
    smoke)   timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    *)       timeout 900 bash -c "$step" ; echo "custom rc=$?" ;;
  esac
  echo "step $step took $(( $(date +%s) - t0 )) s"
done
