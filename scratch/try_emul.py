import sys, time, ctypes as C, numpy as np, pickle
sys.path.insert(0,'.')
from oracle import oracle as O
from bioinformatics_toolkit_b200 import synth
L=C.CDLL('tests/emul/libemul.so'); L.emul_scan.restype=C.c_int64
def emul(mats, thres, dna, strands=2, force_s1=0, bg=O.DEF_BG):
    ncol,off,flat=O._pack_mats(mats)
    cap=1<<22
    om=np.empty(cap,np.int32); os_=np.empty(cap,np.uint8); op=np.empty(cap,np.int64); osc=np.empty(cap)
    st=np.zeros(8,np.int64)
    bg=np.array(bg,float); th=np.array(thres,float); dna=np.ascontiguousarray(dna,np.uint8)
    n=L.emul_scan(C.c_int(len(mats)), ncol.ctypes.data_as(C.c_void_p), flat.ctypes.data_as(C.c_void_p), bg.ctypes.data_as(C.c_void_p),
        th.ctypes.data_as(C.c_void_p), C.c_int(strands), dna.ctypes.data_as(C.c_void_p), C.c_int64(dna.size), C.c_int(force_s1),
        om.ctypes.data_as(C.c_void_p), os_.ctypes.data_as(C.c_void_p), op.ctypes.data_as(C.c_void_p), osc.ctypes.data_as(C.c_void_p), C.c_int64(cap), st.ctypes.data_as(C.c_void_p))
    return n,(om[:n],os_[:n],op[:n],osc[:n]),st
mats=synth.synth_pwms(100,12)
thres=pickle.load(open('scratch/thres_c2.pkl','rb'))
dna=synth.synth_dna(300000, 7, gc=0.41, n_runs=[(1000,500),(200000,3000)])
for fs in (0,2,3):
    t=time.time(); n,h,st=emul(mats,thres,dna,force_s1=fs); te=time.time()-t
    t=time.time(); n2,h2=O.scan_mt(O.DEF_BG,mats,thres,dna,mode=1,nthreads=8,cap=1<<22); to=time.time()-t
    print('force_s1',fs,'emul hits',n,'oracle hits',n2,'cand',st[0],'pwarp(ppm)',st[1],'blocks',st[2],'generic',st[3],'cfg',st[4],'time',round(te,2),round(to,2))
    ok=all(np.array_equal(a,b) for a,b in zip(h,h2))
    print('  identical:',ok)
