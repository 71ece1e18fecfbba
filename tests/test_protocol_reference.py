"""The protocol tables (enhancershoebox_b200/protocol.py) against the reference drivers' OWN source: every
`Group.L.pair_coeff(...)` statement of VisitorGeneMaps/run_single_GENES_STAGES.py and run_single_shoebox.py is
read out of the script's syntax tree together with the `if` conditions that enclose it, the statements are replayed
in file order for a given Python timestep and condition with LAMMPS' wildcard / override rules, and the resulting
(type i, type j) -> (KEY, cut) maps and soft cutoffs must equal what protocol.py hands to the device.  Runs where the
reference checkout is present (this container); skipped on the GPU box."""
import ast
import os
import warnings

import numpy as np
import pytest

from enhancershoebox_b200 import protocol as P

REF = "/root/reference"
DRIVERS = {"genes_stages": os.path.join(REF, "VisitorGeneMaps", "run_single_GENES_STAGES.py"),
           "shoebox": os.path.join(REF, "run_single_shoebox.py")}
pytestmark = pytest.mark.skipif(not os.path.exists(DRIVERS["shoebox"]), reason="reference checkout not present")


def _constants(tree):
    """Module-level `name = <expression of earlier names and numbers>` assignments (sig_*, c_*, t_soft, ...)."""
    env = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            try:
                env[node.targets[0].id] = eval(compile(ast.Expression(node.value), "<ref>", "eval"), {"__builtins__": {}}, dict(env))
            except Exception:
                pass
    return env


def _pair_coeff_statements(path):
    """[(line, [enclosing (test source, taken branch)], argument expressions)] in file order."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)          # the drivers' plot labels hold unescaped backslashes
        tree = ast.parse(open(path).read())
    out = []

    def walk(node, guards):
        for child in ast.iter_child_nodes(node):
            if isinstance(child, ast.If):
                src = ast.unparse(child.test)
                for sub in child.body:
                    walk_stmt(sub, guards + [(src, True)])
                for sub in child.orelse:
                    walk_stmt(sub, guards + [(src, False)])
            else:
                walk_stmt(child, guards)

    def walk_stmt(stmt, guards):
        if isinstance(stmt, ast.If):
            walk(ast.Module(body=[stmt], type_ignores=[]), guards)
            return
        if isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Call) and isinstance(stmt.value.func, ast.Attribute) \
                and stmt.value.func.attr == "pair_coeff":
            out.append((stmt.lineno, list(guards), stmt.value.args))
            return
        walk(stmt, guards)

    walk(tree, [])
    return sorted(out, key=lambda q: q[0]), _constants(tree)


def _replay(driver, condition, i):
    """pair_coeff commands the driver issues at Python timestep i (None: the set-up block before the main loop)."""
    stmts, env = _pair_coeff_statements(DRIVERS[driver])
    env = dict(env, condition=condition, is_actin_simulation=False, print_verbose=0, make_snapshots=0)
    cmds = []
    for line, guards, args in stmts:
        in_loop = any("i" in {n.id for n in ast.walk(ast.parse(src)) if isinstance(n, ast.Name)} for src, _ in guards)
        if (i is None) == in_loop:
            continue
        local = dict(env, i=i)
        try:
            ok = all(bool(eval(src, {"__builtins__": {}}, local)) == taken for src, taken in guards)
        except NameError:
            continue
        if ok:
            cmds.append([eval(compile(ast.Expression(a), "<ref>", "eval"), {"__builtins__": {}}, local) for a in args])
    return cmds


@pytest.mark.parametrize("driver", ["genes_stages", "shoebox"])
def test_soft_cutoffs_are_the_drivers_own(driver):
    cmds = _replay(driver, "Control", None)
    assert len(cmds) == 24 and all(c[2] == 0 for c in cmds)              # pair_coeff I J 0 cutoff (A = 0: fix adapt ramps it)
    want = P._replay([(c[0], c[1], float(c[3])) for c in cmds])
    m = P.soft_cutoffs()
    for (a, b), v in want.items():
        assert m[a, b] == v == m[b, a], (a, b)


@pytest.mark.parametrize("driver,condition", [("genes_stages", "Control"), ("shoebox", "Control"), ("shoebox", "SwinA_nopol"),
                                              ("shoebox", "noActinSer5P"), ("shoebox", "Hexanediol"), ("shoebox", "JQ-1")])
def test_table_maps_are_the_drivers_own(driver, condition):
    def table_cmds(i):
        out = []
        for c in _replay(driver, condition, i):
            name, key = c[2].split()
            assert name == "LJsoft_RSQ.table"
            out.append((c[0], c[1], (key, float(c[3]))))
        return out
    stream_a = table_cmds(P.T_SOFT - 1)                                   # (i+1) == t_soft
    stream_b = table_cmds(P.T_SER5P_ON - 1)                               # (i+1) == t_ser5p_on
    stream_t = table_cmds(P.N_RUNS)                                       # i == NRuns (treatments)
    assert len(stream_a) == (34 if driver == "shoebox" else 27) and len(stream_b) == (8 if driver == "shoebox" else 7)
    assert P._replay(stream_a) == P.table_map(driver, condition, "A")
    assert P._replay(stream_a + stream_b) == P.table_map(driver, condition, "B")
    if condition in ("Hexanediol", "JQ-1"):
        assert stream_t and P._replay(stream_a + stream_b + stream_t) == P.table_map(driver, condition, "T")
    else:
        assert not stream_t
    assert not table_cmds(P.T_SER5P_ON + 3) and not table_cmds(500)      # no other timestep touches the tables


def test_timeline_constants_are_the_drivers_own():
    _, env = _pair_coeff_statements(DRIVERS["shoebox"])
    assert env["t_soft"] == P.T_SOFT and env["t_ser5p_on"] == P.T_SER5P_ON and env["NRuns"] == P.N_RUNS and env["tRun"] == P.T_RUN
    _, env = _pair_coeff_statements(DRIVERS["genes_stages"])
    assert env["n_soft"] == P.T_SOFT and env["t_ser5p_on"] == P.T_SER5P_ON and env["NRuns"] == P.N_RUNS and env["tRun"] == P.T_RUN
