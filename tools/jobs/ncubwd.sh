ncu --set full --clock-control none --import-source on -k regex:"marg_backward_kernel|marg_forward_tail_kernel" -c 2 -f -o gpurun_out/r02z_bwd_tail python tools/profile_driver.py > /dev/null 2>&1
ls -la gpurun_out/r02z_bwd_tail.ncu-rep
