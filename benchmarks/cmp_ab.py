import numpy as np, sys
a=np.load(sys.argv[1]); b=np.load(sys.argv[2])
for k in a.files:
    if k.endswith('_acc'): print(k, a[k], b[k])
    else:
        d=np.abs(a[k]-b[k]); rows=(d.reshape(len(d),-1).max(axis=1)>1e-9).mean()
        print(k, 'max abs diff', d.max(), 'frac rows differing >1e-9', rows)
