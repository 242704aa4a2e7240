# usage: bash tools/profile_kernels.sh <tag> k_name1 k_name2 ...
# One `ncu --set full` capture per kernel (1 launch each, taken after warm-up), exported as text on the box so
# that gpurun_out stays small; plus the launch list of a short bench run.
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 3 -c 1 -o /tmp/prof_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page details > gpurun_out/${tag}_${k}_details.txt 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page source --csv 2>&1 | gzip > gpurun_out/${tag}_${k}_source.csv.gz
done
du -sh gpurun_out
