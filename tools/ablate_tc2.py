import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
code = '''
import sys; sys.argv=["x"]; __file__ = "%s/ablate_tc.py"
exec(open(__file__).read().split("for name, base in")[0])
print("skeleton x-only", run(77 + 256*15, spectral=False), " mainloop(no epi) x-only", run(77, spectral=False), " full", run(1))
''' % (here,)
for nst, nraw in ((4, 8), (3, 8), (2, 8), (4, 3), (4, 2)):
    env = dict(os.environ, PDES_V3_NST=str(nst), PDES_V3_NRAW=str(nraw))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("NST", nst, "NRAW", nraw, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
