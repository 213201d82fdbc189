# ncu evidence of round 2 (one GPU): launch list of the bench command + --set full captures of the kernels added this round
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --quick --blocks 2 > gpurun_out/r02_ncu_bench.log 2>&1; echo launches rc $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dec_persist_kernel -s 2 -c 1 -o gpurun_out/r02_persist -f python tools/trace_persist.py 148 > gpurun_out/r02_ncu_persist.log 2>&1; echo persist rc $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ce_fixup|ce_merge|gemm_tc_kernel<128, 2, 5, 0, 0, 0, 3" -s 3 -c 3 -o gpurun_out/r02_fusedce -f python tools/prof_fused_loss.py > gpurun_out/r02_ncu_fusedce.log 2>&1; echo fusedce rc $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<64, 4, 4, 0, 0, 1, 0, 1>|argmax_filter|argmax_refine|dec_cell" -s 5 -c 5 -o gpurun_out/r02_decode -f python tools/prof_decode.py 4096 3 > gpurun_out/r02_ncu_decode.log 2>&1; echo decode rc $?
for r in persist fusedce decode; do ncu -i gpurun_out/r02_$r.ncu-rep --page details > gpurun_out/r02_${r}_ncu_details.txt 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep | head
