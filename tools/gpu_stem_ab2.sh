# in-slice A/B of the streaming fp32 kernels on one box: packed (current) vs the previous scalar / interleaved version
probe() { python tools/gpu_probe.py n53_m20_sparse1024 --top 60 2>&1 | grep -E "by class|algo=2" | head -n 14 | sed "s/^/$1 /" | cut -c1-150; }
probe packed
cp artensor_b200/csrc/stem.cu /tmp/stem_new.cu; cp gpurun_in_ab/stem_prev.cu artensor_b200/csrc/stem.cu
make -C artensor_b200/csrc -j8 > /dev/null 2>&1
probe prev
cp /tmp/stem_new.cu artensor_b200/csrc/stem.cu; make -C artensor_b200/csrc -j8 > /dev/null 2>&1
probe packed
