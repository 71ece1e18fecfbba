"""The `lammps`-shaped front end (enhancershoebox_b200/lammps_shim.py) driven by the reference driver's own command
stream -- the set-up block of VisitorGeneMaps/run_single_GENES_STAGES.py:533-652 and the main loop's engine calls
(:738-823: soft chunks with `run 200 start 0 stop 2000`, the table switch at (i+1) == n_soft, Ser5P at (i+1) ==
t_ser5p_on, gather_atoms, `set atom ID type T`) -- must leave the same bits behind as the fused batch path
(ShoeboxBatch.set_schedule + run) with the same seed, because it is the same engine behind another front door."""
import os

import numpy as np
import pytest

from enhancershoebox_b200 import datafile, engine as E, protocol as P
from enhancershoebox_b200.lammps_shim import IPyLammps, lammps

pytestmark = pytest.mark.gpu


def _driver_setup(L, path, seed):
    """VisitorGeneMaps/run_single_GENES_STAGES.py:533-652, argument for argument (numbers as the script computes them)."""
    sig_chr, sig_rbp, sig_rnp, sig_ser5p = 1.0, 0.3, 1.0, 0.3
    w = 2 ** (1 / 6)
    c_cc, c_crb, c_crn, c_cs = w, ((sig_chr + sig_rbp) / 2) * w, ((sig_chr + sig_rnp) / 2) * w, ((sig_chr + sig_ser5p) / 2) * w
    c_rbrn, c_rbs, c_rns = ((sig_rbp + sig_rnp) / 2) * w, ((sig_rbp + sig_ser5p) / 2) * w, ((sig_rnp + sig_ser5p) / 2) * w
    L.atom_style('angle'); L.boundary('f', 'f', 'f')
    L.read_data(path, 'extra/special/per/atom', 10, 'extra/bond/per/atom', 10, 'extra/angle/per/atom', 10)
    L.neighbor(0.3, 'multi'); L.neigh_modify('every', 1, 'delay', 10, 'check', 'yes', 'one', 10000, 'page', 100000); L.comm_modify('cutoff', 6)
    L.angle_style('cosine'); L.angle_coeff(1, 100 / 60); L.angle_coeff(2, 33.0)
    L.pair_style('soft', 2 ** (1 / 6))
    for i, j, c in (('*', '*', w), ('1*2', '4', c_crb), ('1*2', '5', 1.5 * c_crn), ('1*2', '6', c_cc), ('1*2', '7', 1.5 * c_cc), ('1*2', '8', 1.0 * c_cs),
                    ('4', '4', sig_rbp * c_cc), ('4', '5', c_rbrn), ('4', '6*7', c_crb), ('4', '8', c_rbs), ('4', '9*11', c_crb),
                    ('5', '5', sig_rnp * c_cc), ('5', '6', 1.5 * c_crn), ('5', '7', c_crn), ('5', '8', c_rns), ('5', '9*11', 1.5 * c_crn),
                    ('6', '7', 1.5 * c_cc), ('6', '8', c_cs), ('6', '9*11', c_cc), ('7', '7', 1.5 * c_cc), ('7', '8', 1.0 * c_cs), ('7', '9*11', 1.5 * c_cc),
                    ('8', '8', sig_ser5p * c_cc), ('8', '9*11', c_cs)):
        L.pair_coeff(i, j, 0, c)
    L.variable('prefactor equal ramp(0,60)'); L.fix('s1 all adapt 1 pair soft a * * v_prefactor')
    L.bond_style('harmonic'); L.bond_coeff(1, 90, 1); L.bond_coeff(2, 90, 0.3)
    L.variable('chromatinr0 equal ramp(0.033,1.0)'); L.fix('s2 all adapt 10 bond harmonic r0 1 v_chromatinr0')
    L.fix(1, 'all', 'nve'); L.fix(2, 'all', 'langevin', 1, 1, 0.5, seed)
    L.fix('wallhi all wall/harmonic xlo EDGE 100 0.0 3.0 xhi EDGE 100 0.0 3.0 ylo EDGE 100 0.0 3.0 yhi EDGE 100 0.0 3.0 zlo EDGE 100 0.0 3.0 zhi EDGE 100 0.0 3.0')
    L.group('Anchors', 'id', 1, 123, 124, 370); L.fix('freezeAnchors', 'Anchors', 'setforce', 0, 0, 0)
    L.timestep(0.005)


def test_the_drivers_command_stream_gives_the_batch_engines_bits(tmp_path):
    d = datafile.make_shoebox(11, 3, False, seed=3)
    path = os.path.join(tmp_path, "IC.data")
    datafile.write_data(path, d)
    d = datafile.read_data(path)                                          # (free species are written with 6 decimals)
    n_chunks, flip_at, flip_atom = 17, 15, 700
    lmp = lammps()
    L = IPyLammps(ptr=lmp)
    _driver_setup(L, path, 4711)
    assert L.system.natoms == 1290 and L.system.xlo == d.box_lo[0]
    table_a = P._table_stream_phase_a("genes_stages", "Control")
    table_b = P._table_stream_phase_b("genes_stages", "Control")
    for i in range(n_chunks):
        if (i + 1) == 10:
            L.unfix('s1'); L.unfix('s2'); L.pair_style('table lookup 30000')
            for a, b, (key, cut) in table_a:
                L.pair_coeff(a, b, 'LJsoft_RSQ.table ' + key, cut)
        if (i + 1) == 15:
            for a, b, (key, cut) in table_b:
                L.pair_coeff(a, b, 'LJsoft_RSQ.table ' + key, cut)
        if (i + 1) < 10:
            L.run(200, 'start 0 stop', 10 * 200)
        else:
            L.run(200)
        x = np.array(lmp.gather_atoms('x', 1, 3)).reshape(-1, 3)
        t = np.array(lmp.gather_atoms('type', 0, 1))
        if i == flip_at:
            L.set('atom', flip_atom + 1, 'type', 5)                       # an RBP becomes an RNP (run_single_shoebox.py:1469)
    # the same protocol through the batch engine: chunk by chunk, same seed, the same type change at the same point
    eng = E.ShoeboxBatch(d, n_replicas=1, seeds=[4711])
    eng.set_schedule(n_chunks, observe=False)
    eng.run(0, flip_at + 1)
    xb, vb, tb = eng.state()
    tb[0, flip_atom] = 5
    eng.upload_state(xb, vb, tb)
    eng.run(flip_at + 1, n_chunks - flip_at - 1)
    xb, vb, tb = eng.state()
    assert np.array_equal(x, xb[0]) and np.array_equal(t, tb[0]) and t[flip_atom] == 5
    assert np.abs(x - d.x).max() > 1.0
    with pytest.raises(Exception):
        L.minimize(1e-4, 1e-6, 100, 1000)                                 # not in the drivers' command set
    lmp.close()
    eng.close()
