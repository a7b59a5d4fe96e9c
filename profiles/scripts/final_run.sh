#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/r2_final_c3.log 2> gpurun_out/r2_final_c3.err
for c in 0 1 2 4; do timeout 400 python bench.py --config $c > gpurun_out/r2_final_c$c.log 2> gpurun_out/r2_final_c$c.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv -k regex:"k_chol|k_trinv|k_solve3|k_post_fft|k_sample" -s 120 -c 90 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --substreams 1 > gpurun_out/r2_final_ncu.log 2>&1
cat gpurun_out/r2_final_pytest.log gpurun_out/r2_final_smoke.log | tail -5
