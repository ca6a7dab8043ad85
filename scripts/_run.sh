set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_gpu_tests.log 2>&1; tail -3 gpurun_out/r2v_gpu_tests.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -c 300 gpurun_out/r2v_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err; tail -c 300 gpurun_out/r2v_bench_ref.json
python scripts/gen_timing.py > gpurun_out/r2v_gen_timing.log 2>&1; cat gpurun_out/r2v_gen_timing.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline --no-strong > gpurun_out/r2v_ncu_bench.log 2>&1
mkdir -p /tmp/rep
SPECS=""
for prec in f64 f64fused f32 f32fast; do
  ncu -k regex:decode_resident --launch-skip 1 -c 1 --set full --import-source on --clock-control none -o /tmp/rep/r2v_resident_$prec -f python scripts/profile_point.py $prec 0.09 2960 > gpurun_out/r2v_ncu_$prec.log 2>&1
  FI=$(grep -o "frame-iterations=[0-9]*" gpurun_out/r2v_ncu_$prec.log | tail -1 | cut -d= -f2)
  BPE=16; case $prec in f64*) BPE=32;; esac
  python scripts/ncu_summary.py /tmp/rep/r2v_resident_$prec.ncu-rep 24 > gpurun_out/r02_ncu_resident_${prec}_q009_2960frames.txt 2>&1
  SPECS="$SPECS $prec=/tmp/rep/r2v_resident_$prec.ncu-rep:$FI:$BPE:30720:2960_reference_frames_at_QBER_0.09_(scripts/profile_point.py)"
done
ncu -k regex:stream_bit_kernel --launch-skip 24 -c 1 --set full --import-source on --clock-control none -o /tmp/rep/r2v_stream_bit_f64 -f python scripts/stream_probe.py 100000 51080 9472 0.10 20 f64 > gpurun_out/r2v_ncu_sbit.log 2>&1
python scripts/ncu_summary.py /tmp/rep/r2v_stream_bit_f64.ncu-rep 24 > gpurun_out/r02_ncu_stream_bit_f64_n100k.txt 2>&1
ncu -k regex:stream_check_kernel --launch-skip 24 -c 1 --set full --import-source on --clock-control none -o /tmp/rep/r2v_stream_check_f64 -f python scripts/stream_probe.py 100000 51080 9472 0.10 20 f64 > gpurun_out/r2v_ncu_scheck.log 2>&1
python scripts/ncu_summary.py /tmp/rep/r2v_stream_check_f64.ncu-rep 24 > gpurun_out/r02_ncu_stream_check_f64_n100k.txt 2>&1
SPECS="$SPECS stream_n100k_f64_check=/tmp/rep/r2v_stream_check_f64.ncu-rep:9472:16:300000:one_check_pass_over_9472_frames_of_the_N=100000_code stream_n100k_f64_bit=/tmp/rep/r2v_stream_bit_f64.ncu-rep:9472:16:300000:one_bit_pass_over_9472_frames_of_the_N=100000_code"
python scripts/make_traffic.py $SPECS > gpurun_out/r2v_make_traffic.log 2>&1; cp profiles/traffic.json gpurun_out/r2v_traffic.json
bash scripts/bounds_check_build.sh run > gpurun_out/r2v_bounds.log 2>&1; echo "bounds rc=$?"; tail -3 gpurun_out/r2v_bounds.log
du -sh gpurun_out
