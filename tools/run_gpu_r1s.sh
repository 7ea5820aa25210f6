mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  for c in n53_m20_sparse1024 n30_sparse10000; do
    env "$@" timeout -s KILL 400 python tools/gpu_probe.py $c --top 60 --tag _$tag > gpurun_out/p_${c}_$tag.log 2>&1
    echo "$tag $c: $(sed -n 3,4p gpurun_out/p_${c}_$tag.log | tr '\n' ' ')"
  done
}
run old TNC_STEM_NO_BULK=1
run c2 TNC_STEM_BULK_CTAS=2
run c3 TNC_STEM_BULK_CTAS=3
run old2 TNC_STEM_NO_BULK=1
