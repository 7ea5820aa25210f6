# A/B of the packed (fma.f32x2) and scalar streaming fp32 kernels on the same box: isolated steps of the n53 tree's shapes
run() { for sh in "25 3 3" "26 2 4" "26 3 2" "24 3 3" "24 4 4"; do python tools/one_step.py $sh --algo stem --shuffle --reps 3 2>&1 | tail -n 1 | sed "s/^/$1 $sh: /"; done; }
run packed
touch artensor_b200/csrc/stem.cu; make -C artensor_b200/csrc EXTRA=-DTNC_STEM_SCALAR -j8 > /dev/null 2>&1
run scalar
