"""Per-kernel DRAM traffic and time of ONE forward from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--csv` launch list (captured inside the NVTX range "measured"):  python tools/dram_summary.py <csv> <out.json> [ir_split]
ir_split = N: the first N modconv_gemm_kernel launches are reported as `conv_gemm_ir` (IR-SE50 trunk + heads run before the decoder)."""
import collections
import csv
import json
import re
import sys

path, out = sys.argv[1], sys.argv[2]
ir_split = int(sys.argv[3]) if len(sys.argv) > 3 else 0
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith('==')]))


def val(row):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit'].lower()
    scale = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3,
             'nsecond': 1e-3}
    return v * scale.get(u, 1.0)


launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r['ID'], {'name': re.sub(r'^void ', '', re.sub(r'\(.*', '', r['Kernel Name']))})
    d[r['Metric Name']] = val(r)
agg = collections.OrderedDict()
n_gemm = 0
for d in launch.values():
    name = d['name']
    if 'modconv_gemm_kernel' in name:
        n_gemm += 1
        if ir_split:
            name = 'conv_gemm_ir: ' + name if n_gemm <= ir_split else 'conv_gemm: ' + name
    a = agg.setdefault(name[:90], {'launches': 0, 'us': 0.0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0})
    a['launches'] += 1
    a['us'] += d.get('gpu__time_duration.sum', 0.0)
    a['dram_read_bytes'] += d.get('dram__bytes_read.sum', 0.0)
    a['dram_write_bytes'] += d.get('dram__bytes_write.sum', 0.0)
tot = sum(a['us'] for a in agg.values())
res = {'source': path.split('/')[-1], 'launches': len(launch), 'sum_us_cold_serialised': tot,
       'kernels': sorted(({'kernel': k, **v, 'share_of_time': v['us'] / tot} for k, v in agg.items()), key=lambda x: -x['us'])}
json.dump(res, open(out, 'w'), indent=1)
for k in res['kernels'][:25]:
    print(f"{k['us']:10.1f} us {k['launches']:5d}  rd {k['dram_read_bytes'] / 1e6:9.1f} MB  wr {k['dram_write_bytes'] / 1e6:9.1f} MB  {k['share_of_time']:.3f}  {k['kernel']}")
