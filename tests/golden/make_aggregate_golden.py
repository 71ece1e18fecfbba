"""Runs the REFERENCE aggregator VisitorGeneMaps/dist_genestages_grouped.py UNMODIFIED on a synthetic
geneTrack tree and stores inputs + its CSV output as tests/golden/aggregate_golden.npz.

The script's plotting imports (matplotlib, seaborn; never used) are stubbed, and os.listdir is
wrapped so that run folders come back in numeric order (the script takes them in listdir order,
VisitorGeneMaps/dist_genestages_grouped.py:63-64).  Needs /root/reference (build container only)."""
import os, runpy, sys, tempfile, types
import numpy as np

REF = '/root/reference/VisitorGeneMaps/dist_genestages_grouped.py'
REF_PCT = '/root/reference/VisitorGeneMaps/dist_genestages_grouped_5percentile_10xaveraging.py'
rng = np.random.default_rng(0)


def synth_run(n_frames, p_on, rng):
    """geneTrack columns for one run: state script 0 (150 frames) -> 1, activation windows of 50."""
    t = (np.arange(n_frames) + 1) * 6.0
    probe = rng.uniform(-600, 600, (n_frames, 3))
    d_rp = rng.uniform(50, 900, n_frames)
    d_rg = rng.uniform(50, 900, n_frames)
    s5p = rng.integers(0, 120, n_frames)
    s5g = rng.integers(0, 120, n_frames)
    state = np.zeros(n_frames, int)
    k = 150
    while k < n_frames:
        if rng.random() < p_on:
            state[k:k + 50] = 2
            if k + 50 < n_frames:
                state[k + 50] = 0
            k += 51
        else:
            state[k] = 1
            k += 1
    return np.column_stack([t, probe, d_rp, d_rg, s5p, s5g, state]).astype(float)


def main():
    conds = [((3, 80, 30), 10, 2000, 0.004), ((1, 10, 1), 10, 2000, 0.00005), ((2, 50, 10), 5, 10000, 0.002)]
    tmp = tempfile.mkdtemp()
    frames = {}
    for (p, t, a), n_runs, n_frames, p_on in conds:
        root = os.path.join(tmp, 'box11', f'Control_Promoter{p}_Threshold{t}_Act{a}')
        os.makedirs(root)
        open(os.path.join(root, 'dist_active.txt'), 'w').write('1.0\n')
        arr = np.stack([synth_run(n_frames, p_on, rng) for _ in range(n_runs)])
        frames[(p, t, a)] = arr
        for r in range(n_runs):
            os.makedirs(os.path.join(root, f'run{r + 1}'))
            with open(os.path.join(root, f'run{r + 1}', 'geneTrack.txt'), 'w') as f:
                for row in arr[r]:
                    # the drivers write str(float) for the measured columns and str(int) for counts/state
                    cols = [str(float(v)) for v in row[:6]] + [str(int(row[6])), str(int(row[7])), str(int(row[8]))]
                    f.write(','.join(cols) + '\n')
    for name in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):
        sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    real_listdir = os.listdir

    def ordered_listdir(path):
        items = real_listdir(path)
        return sorted(items, key=lambda s: (0, int(s[3:])) if s.startswith('run') and s[3:].isdigit() else (1, s))
    os.listdir = ordered_listdir
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        runpy.run_path(REF, run_name='__main__')
        runpy.run_path(REF_PCT, run_name='__main__')          # the 5-percentile / 10x-averaged variant, same tree
    finally:
        os.chdir(cwd)
        os.listdir = real_listdir
    csv = open(os.path.join(tmp, 'summary_contact_grouped_Thresholds10-100_test.txt')).read()
    csv_pct = open(os.path.join(tmp, 'summary_contact_grouped_Thresholds10-100_5percentile_10xaveraged.txt')).read()
    out = {'csv': np.array(csv), 'csv_pct': np.array(csv_pct)}
    for (p, t, a), arr in frames.items():   # only the columns the aggregator reads are kept (size)
        out[f'drp_{p}_{t}_{a}'] = arr[:, :, 4]
        out[f'drg_{p}_{t}_{a}'] = arr[:, :, 5]
        out[f's5p_{p}_{t}_{a}'] = arr[:, :, 6].astype(np.int16)
        out[f'state_{p}_{t}_{a}'] = arr[:, :, 8].astype(np.int8)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'aggregate_golden.npz'), **out)
    print(csv)
    print(csv_pct)


if __name__ == '__main__':
    main()
