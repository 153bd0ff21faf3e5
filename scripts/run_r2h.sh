# round-2 evidence run (under gpurun): default bench line, reference arm, configs[3] record, ncu passes
T=${TAG:-r2h}
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_${T}_reference.json 2>&1; echo "ref rc=$?"
python bench.py --size 16384 --scales 8 --warps 10 --dx 9 --dy -6 --shear-px 0.5 --seed 13 --margin 64 --coarse 16 \
   --steps 2 --warmup 3 --no-cpu --no-parity --stack-pairs 0 --volume-pairs 0 > gpurun_out/bench_${T}_config3_16384.json 2> gpurun_out/bench_${T}_config3.err; echo "c3 rc=$?"
python bench.py --size 2048 --scales 5 --warps 5 --shear-px 4.096 --steps 20 --warmup 5 --no-cpu --stack-pairs 0 --volume-pairs 0 > gpurun_out/bench_${T}_config0_2048.json 2> gpurun_out/bench_${T}_config0.err; echo "c0 rc=$?"
TAG=$T bash scripts/ncu_r2.sh
