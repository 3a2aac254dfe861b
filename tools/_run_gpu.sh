mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_v2.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2_launches_bench_v2.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_bench_v2.csv')) if len(r) > 10]
hdr = rows[0]; ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(',', '')); u = r[iu]
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1)
    tot[r[ik][:60]] += ns; n[r[ik][:60]] += 1
s = sum(tot.values())
with open('gpurun_out/r2_launch_shares_v2.csv', 'w') as f:
    f.write('# kernel, launches, total ns, share   (python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline under ncu --metrics gpu__time_duration.sum; last build of round 2)\n')
    for k, v in tot.most_common(8):
        f.write('"%s",%d,%d,%.5f\n' % (k, n[k], v, v / s))
print(open('gpurun_out/r2_launch_shares_v2.csv').read())
PY
