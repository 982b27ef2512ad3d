mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 6000 gpurun_out/r2i_bench.json; tail -5 gpurun_out/r2i_bench.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err; cat gpurun_out/r2i_ref.json; tail -5 gpurun_out/r2i_ref.err
