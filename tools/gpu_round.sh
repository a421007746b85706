#!/bin/bash
# One GPU-box session: parity tests, the default bench line, the ncu launch list + one full capture of the resident
# kernel, and a racecheck / memcheck pass.  Everything lands under gpurun_out/<tag>_*.
tag=${1:-r2}; shift
what=${@:-tests bench launches ncu sanitize}
mkdir -p gpurun_out
cp maro_b200/libmaro_b200.so gpurun_out/${tag}_lib.so  # the exact binary the captures refer to (tools/ncu_by_line.py)
for w in $what; do case $w in
tests)    timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log;;
bench)    timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; head -c 1500 gpurun_out/${tag}_bench.json;;
ref)      timeout 600 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/${tag}_bench_reference.json;;
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 640 --warmup 64 --skip-extras --skip-e2e --cpu-seconds 0.5 > gpurun_out/${tag}_launches_bench.log 2>&1; echo "launches rc=$?";;
ncu)      timeout 900 ncu --set full --clock-control none --import-source on -k regex:cim_resident -c 3 -f -o gpurun_out/${tag}_resident_1k python tools/ncu_target.py toy.4p_ssdd_l0.0 1024 1000 64 4 rollout > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?";;
sanitize) timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/ncu_target.py toy.4p_ssdd_l0.0 64 200 48 2 rollout > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/${tag}_racecheck.log
          timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/ncu_target.py toy.4p_ssdd_l0.0 64 200 48 2 step > gpurun_out/${tag}_racecheck_step.log 2>&1; echo "racecheck(step) rc=$?"; tail -3 gpurun_out/${tag}_racecheck_step.log
          timeout 600 compute-sanitizer --tool memcheck python tools/ncu_target.py toy.4p_ssdd_l0.0 64 200 48 2 rollout > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${tag}_memcheck.log;;
esac; done
