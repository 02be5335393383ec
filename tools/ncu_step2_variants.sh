# ncu --set full of k_step2 variants at 1024 x 1024 x 512: tools/ncu_step2_variants.sh "<wy> <stages> <kchunk> <persistent> <window>" ...
for cfg in "$@"; do
  set -- $cfg
  tag=k4_wy$1_s$2_kc$3_p$4_w$5
  ncu --set full --clock-control none -k regex:k_step2 -s 2 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 4 --warmup 4 --nz 512 --no-e2e --no-cpu --no-selfcheck --opt kernel=4 --opt warps_y=$1 --opt stages=$2 --opt kchunk=$3 --opt persistent=$4 --opt window=$5 > gpurun_out/ncu_$tag.log 2>&1
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$tag.ncu-rep --page details > gpurun_out/prof_${tag}_details.txt 2>/dev/null
  echo $tag; python tools/ncu_traffic.py gpurun_out/prof_${tag}_raw.csv --cells 536870912
done
