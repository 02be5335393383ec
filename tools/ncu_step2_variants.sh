for cfg in "8 3 32 1" "8 3 64 1" "8 3 64 4" "16 3 32 1"; do
  set -- $cfg
  tag=k4_wy$1_s$2_kc$3_b$4
  ncu --set full --clock-control none -k regex:k_step2 -s 2 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 4 --warmup 4 --nz 512 --no-e2e --no-cpu --no-selfcheck --opt kernel=4 --opt warps_y=$1 --opt stages=$2 --opt kchunk=$3 --opt band=$4 > gpurun_out/ncu_$tag.log 2>&1
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
  echo $tag; python tools/ncu_traffic.py gpurun_out/prof_${tag}_raw.csv --cells 536870912
done
