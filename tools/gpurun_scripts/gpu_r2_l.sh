#!/bin/bash
# ncu launch lists of the final build: (1) whole bench command at the configs[1] size, (2) the step kernels only at n = 50k
mkdir -p gpurun_out
CMD1="python bench.py --n 10768 --steps 5 --warmup 3 --no-e2e --no-cpu --no-elbo --no-lazy"
timeout 600 $CMD1 > gpurun_out/ncu_final_plain1.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_n10k_final.csv $CMD1 > gpurun_out/ncu_final_list1.log 2>&1
echo "list1 exit $?"; grep -c '^"' gpurun_out/launches_n10k_final.csv
CMD2="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-elbo --no-lazy"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'downdate|score|select|segments|unpack' -c 400 --csv --log-file gpurun_out/launches_n50k_steps_final.csv $CMD2 > gpurun_out/ncu_final_list2.log 2>&1
echo "list2 exit $?"; grep -c '^"' gpurun_out/launches_n50k_steps_final.csv
