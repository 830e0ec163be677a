#!/bin/bash
# Round 2, call W (1 GPU): the round-end sequence as the driver runs it -- whole GPU suite, smoke(), both bench arms --
# then the ncu launch list of the bench command and one --set full capture of the int8 tcgen05 GEMM (plain runs first).
mkdir -p gpurun_out/r02w
O=gpurun_out/r02w
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider --durations=6 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $O/rc.txt
tail -12 $O/pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/rc.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" | tee -a $O/rc.txt
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
echo "reference rc=$?" | tee -a $O/rc.txt
if grep -q "bench rc=0" $O/rc.txt; then
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-elbo --no-e2e-pageable > $O/ncu_bench.log 2>&1
echo "ncu list rc=$?" | tee -a $O/rc.txt
fi
timeout 200 python tools/emulated_gemm_bench.py 8192 8 > $O/emu_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:emu_gemm_resident -c 1 -f -o $O/emu_prof \
    python tools/emulated_gemm_bench.py 8192 8 > $O/emu_ncu.log 2>&1
echo "ncu full rc=$?" | tee -a $O/rc.txt
tail -1 $O/emu_plain.log | cut -c1-400
ls -la $O | cut -c20-120
python - <<'PY'
import json
for name in ("bench_n1", "bench_ref"):
    try:
        d = json.loads(open("gpurun_out/r02w/%s.json" % name).read().strip().splitlines()[-1])
    except Exception as e:
        print(name, "unreadable", e); continue
    print("==", name, d.get("value"), d.get("unit"), "ms/step", d.get("ms_per_step"), "launches", d.get("gpu_launches"))
    print("roofline", json.dumps(d.get("roofline"))[:500])
    print("e2e", json.dumps(d.get("e2e"))[:900])
    print("cpu", json.dumps(d.get("cpu_baseline"))[:500])
    print("parity", json.dumps(d.get("parity"))[:300], "clocks", json.dumps(d.get("clocks")))
    el = d.get("elbo") or {}
    print("elbo", el.get("value"), el.get("ms_per_step"))
PY
