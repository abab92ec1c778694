timeout 120 python tools/time_kernels.py c2 10 warp,fused
for v in p1c4 p1c3 p0c3 p1c2; do APAP_B200_LIB=cvx_proj_b200/lab/$v.so timeout 120 python tools/time_kernels.py c2 10 warp,fused; done
timeout 120 python tools/time_kernels.py c3 5 warp,fused
for v in p1c4 p1c3 p0c3 p1c2; do APAP_B200_LIB=cvx_proj_b200/lab/$v.so timeout 120 python tools/time_kernels.py c3 5 warp,fused; done
