#!/bin/bash
# One parameterised GPU script (run under gpurun): tools/gpu.sh <what> [...]
#   tests            pytest -m gpu (log: gpurun_out/pytest_gpu.log)
#   smoke            __graft_entry__.smoke()
#   bench [args]     python bench.py args           -> gpurun_out/bench.json
#   benchn N [args]  torchrun bench.py --gpus N ... -> gpurun_out/bench_nN.json
#   ncu [args]       launch list + one --set full capture of the top kernel of `bench.py args`
#   ref [args]       python bench.py --impl reference args
# several commands can be chained with "--":  tools/gpu.sh smoke -- tests -- bench --steps 20
mkdir -p gpurun_out
run_one() {
    what=$1; shift
    case "$what" in
    smoke) python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    tests)
        timeout 1500 python -m pytest tests -m gpu -q -x "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
        tail -15 gpurun_out/pytest_gpu.log ;;
    bench)
        tag=${BENCH_TAG:-bench}
        timeout 900 python bench.py "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err; echo "bench rc=$?"
        tail -3 gpurun_out/$tag.err; python tools/show_bench.py gpurun_out/$tag.json ;;
    benchn)
        n=$1; shift; tag=${BENCH_TAG:-bench_n$n}
        timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
            bench.py --gpus $n "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err; echo "bench n=$n rc=$?"
        tail -3 gpurun_out/$tag.err; python tools/show_bench.py gpurun_out/$tag.json ;;
    ref)
        timeout 900 python bench.py --impl reference "$@" > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
        cut -c1-600 gpurun_out/bench_reference.json ;;
    ncu)
        tag=${NCU_TAG:-default}
        python bench.py "$@" > gpurun_out/ncu_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$tag.log; return; }
        ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
            python bench.py "$@" > gpurun_out/ncu_list_$tag.log 2>&1; echo "ncu list rc=$?"
        ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-k_step_fused}" -s ${NCU_SKIP:-3} -c 1 \
            -f -o gpurun_out/prof_$tag python bench.py "$@" > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
        ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
        ncu -i gpurun_out/prof_$tag.ncu-rep --page details > gpurun_out/prof_${tag}_details.txt 2>/dev/null
        python tools/ncu_traffic.py gpurun_out/prof_${tag}_raw.csv ;;
    *) echo "unknown command $what" ;;
    esac
}
args=()
for a in "$@"; do
    if [ "$a" == "--" ]; then run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_one "${args[@]}"
