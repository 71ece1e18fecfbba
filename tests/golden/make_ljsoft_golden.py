"""Runs the REFERENCE table generator (ljsoft_potential.py, functions only -- the module-level code
that writes the 57 MB file is cut off at `filename=`) for all 55 sections and records, per section,
the SHA-256 of the exact text it writes plus a few rows.  Output: tests/golden/ljsoft_sections.json.
Needs /root/reference (run in the build container only)."""
import hashlib, io, json, os, re, sys

REF = '/root/reference/ljsoft_potential.py'
src = open(REF).read().split('filename="LJsoft_RSQ.table"')
head, tail = src[0].replace('import matplotlib.pyplot as plt', ''), src[1]
ns = {}
exec(compile(head, REF, 'exec'), ns)
# parameters of every write_section call, evaluated with the reference's own variable definitions
prm_src = tail.split('# eps, s, a, b')[0]
exec(compile('\n'.join(l for l in prm_src.split('\n') if not l.startswith('myfile')), REF, 'exec'), ns)
calls = re.findall(r"^(write_section(?:_longrange)?)\('([A-Z0-9_]+)', N, dr, rmax, myfile, lam, n, alpha, (.*?)\)\s*(?:#.*)?$", tail, re.M)
out = {}
for fn, key, args in calls:
    eps0, s0, a, b = eval('(' + args + ')', ns)
    buf = io.StringIO()
    ns[fn](key, ns['N'], ns['dr'], ns['rmax'], buf, ns['lam'], ns['n'], ns['alpha'], eps0, s0, a, b)
    text = buf.getvalue()
    lines = text.split('\n')
    out[key] = dict(sha256=hashlib.sha256(text.encode()).hexdigest(), n_bytes=len(text), header=lines[1],
                    row1=lines[3], row4000=lines[3 + 3999], row30000=lines[3 + 29999],
                    eps0=float(eps0), s0=float(s0), a=float(a), b=float(b), longrange=fn.endswith('longrange'))
    print(key, out[key]['sha256'][:12], len(text), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ljsoft_sections.json'), 'w'), indent=1)
print(len(out), 'sections')
