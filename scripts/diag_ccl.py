import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, scipy.ndimage as ndi
from test_label_gpu import _ccl
for shape,density,full in [((200,333),0.4,True),((200,333),0.55,False),((37,45,70),0.3,False),((64,64),0.4,True),((40,32),0.4,True),((40,31),0.4,True),((40,33),0.4,True)]:
    rng=np.random.default_rng(1)
    mask=rng.random(shape)<density
    ref,nref=ndi.label(mask,structure=np.ones((3,)*mask.ndim,bool) if full else None)
    got,ngot=_ccl(mask,full)
    same_support=np.array_equal(got>0,ref>0)
    # partition equivalence
    pairs=np.unique(np.stack([got[mask],ref[mask]],1),axis=0)
    one2one = len(np.unique(pairs[:,0]))==len(pairs)==len(np.unique(pairs[:,1]))
    print(shape,density,full,"n",ngot,nref,"support",same_support,"partition same",one2one,"exact",np.array_equal(got,ref), "max",got.max())
    if one2one and not np.array_equal(got,ref):
        bad=np.argwhere(got!=ref)[:3]; print("  first diffs",bad.tolist(),[ (int(got[tuple(b)]),int(ref[tuple(b)])) for b in bad])
