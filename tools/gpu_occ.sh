export PB200_LIB=$PWD/pluto_sirocco_b200/lib/libplutob200_hot.so
for pad in 0 70000 40000; do
PB200_SMEM_PAD=$pad timeout 300 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu > /tmp/o.json 2>/tmp/o.err; python -c "
import json; d=json.load(open('/tmp/o.json')); print('pad',$pad, round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()})"
done
