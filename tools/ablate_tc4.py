import os, sys
__file__ = os.path.join(os.path.dirname(os.path.abspath(sys.argv[0])), "ablate_tc.py")
exec(open(__file__).read().split("for name, base in")[0])
for mask in (15, 31, 1, 17):
    print(f"no epilogue mask {mask:2d}: x-only {run(77 + 256 * mask, spectral=False):6.1f} us")
