timeout 120 python tools/time_kernels.py c2 10 gram
for v in poly1 poly3 poly4; do APAP_B200_LIB=cvx_proj_b200/lab/$v.so timeout 120 python tools/time_kernels.py c2 10 gram; done
timeout 120 python tools/time_kernels.py c3 5 gram
for v in poly1 poly3; do APAP_B200_LIB=cvx_proj_b200/lab/$v.so timeout 120 python tools/time_kernels.py c3 5 gram; done
