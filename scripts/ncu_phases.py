"""Sum an ncu source page (cuda,sass CSV) per PHASE of the step kernel: the phases are line ranges of esb_kernels.cu found
from anchor strings in the source itself, so the table follows the code.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python scripts/ncu_phases.py src.csv [esb_kernels.cu]"""
import csv, collections, os, sys
src_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), '..', 'enhancershoebox_b200', 'csrc', 'esb_kernels.cu')
src = open(src_path).read().split('\n')
ANCHORS = [  # (phase, first line containing this string); a phase runs to the next anchor
    ('head', 'namespace {'),
    ('build:sort', 'void slot_sort('),
    ('build:sort', 'void build_lists('),
    ('build:batch_setup', 'batches are handed out dynamically'),
    ('build:spans', "this lane's span: (bead, class, row)"),
    ('build:candidates', 'the concatenated spans, 32 (bead, candidate) pairs'),
    ('build:descriptors', '__syncwarp();\n        if (lane < NB)'),
    ('build:tail', 'force order: descriptors sorted by decreasing list length'),
    ('bonded', '// Bonded forces.  Like the pair forces'),
    ('bonded', 'void bonded_forces('),
    ('force:setup+items', 'void force_phase('),
    ('force:pair_body', 'auto pair_body = [&]'),
    ('force:loop', 'constexpr int DEPTH = ESB_PREFETCH;'),
    ('force:setup+items', 'for (int d = ESB_G >> 1; d >= 1; d >>= 1) {'),
    ('atom', 'int post_force('),
    ('observe', 'double ramp(const double v0'),
    ('runkernel', 'esb_run_kernel(const __grid_constant__ KArgs a'),
    ('hooks', 'esb_forces_kernel(const __grid_constant__ KArgs a'),
]
text = '\n'.join(src)
bounds = []
for ph, key in ANCHORS:
    pos = text.find(key)
    if pos < 0:
        continue
    bounds.append((text.count('\n', 0, pos) + 1, ph))
bounds.sort()
def phase_of(fn, line):
    if 'esb_kernels' not in fn:
        return 'other:' + fn
    ph = 'head'
    for l0, p in bounds:
        if line >= l0: ph = p
        else: break
    return ph
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != '-': continue
    try: n = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[4:], r[4:]))
    smp = float(d['# Samples'] or 0); ins = float(d['Instructions Executed'] or 0); thr = float(d['Thread Instructions Executed'] or 0)
    a = agg[phase_of(cur, n)]
    a[0] += smp; a[1] += ins; a[2] += thr
    tot_s += smp; tot_i += ins
for ph, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{ph:40s} smp {100*s/tot_s:6.2f}%  inst {100*i/tot_i:6.2f}%  thr/inst {t/max(i,1):5.1f}')
print('total samples', tot_s, 'total warp instructions', tot_i)
