"""Bit-level model of the exact parallel evaluation of the float recurrence s <- fl(s + t_k), t_k >= 0, used by the
ordered-sum kernels of the demapper (sdr_receiver_dvb_t2_b200/csrc/demap.cu): ordered_sum() is the pair scan a warp runs
over one chunk (sum_exact_warp), stitched_sum() the whole scheme -- tree-order chunk sums predict the binade of the running
sum, every chunk is added up as ONE integer for that binade (demap_sum_chunks_kernel), a serial walk over the chunks takes
the integer where the prediction holds and falls back to the exact scan where it does not (demap_sum_stitch_kernel).  Run
it to check the algorithm against a serial sum."""
import numpy as np
BIG=1<<26
def elem(Es, tb):
    if tb==0: return 0,0,0
    Et=(tb>>23)&0xff; mt=(tb&0x7fffff)|(0x800000 if Et else 0)
    if Et==0: Et=1
    sh=Es-Et
    if sh<0: return 1<<24,0,0
    if sh==0: return mt,0,0
    if sh<=24:
        q=mt>>sh; rem=mt&((1<<sh)-1); half=1<<(sh-1)
        return q,int(rem>half),int(rem==half)
    return 0,0,0
def sat(x): return min(x,BIG)
def compose(A,B):  # A first
    return [sat(A[p]+B[p^(A[p]&1)]) for p in (0,1)]
def fbits(x): return int(np.float32(x).view(np.uint32))
def frombits(b): return np.uint32(b).view(np.float32)
def ordered_sum(t, CH=64, E=4):
    n=len(t); s=np.float32(0); base=0; passes=0
    while base<n:
        sb=fbits(s); Es=(sb>>23)&0xff
        if Es==0:   # zero/denormal: serial step
            s=np.float32(s+t[base]); base+=1; continue
        S0=(sb&0x7fffff)|0x800000
        m=min(CH,n-base); passes+=1
        nt=(m+E-1)//E
        el=[elem(Es,fbits(t[base+k])) for k in range(m)]
        loc=[]
        for th in range(nt):
            A=[0,0]
            for k in range(th*E,min(m,th*E+E)):
                q,gt,tie=el[k]
                for v in (0,1):
                    cp=v^(A[v]&1)
                    A[v]=sat(A[v]+q+gt+(tie&((cp+q)&1)))
            loc.append(A)
        # exclusive scan
        pref=[[0,0]]
        for th in range(nt-1): pref.append(compose(pref[-1],loc[th]))
        cross=None; Send=None
        for th in range(nt):
            S=S0+pref[th][S0&1]
            if S>=1<<24:
                cross=th*E if cross is None else cross; break
            done=False
            for k in range(th*E,min(m,th*E+E)):
                q,gt,tie=el[k]
                Sn=S+q+gt+(tie&((S+q)&1))
                if Sn>=1<<24:
                    cross=(k,S); done=True; break
                S=Sn
            if done: break
            Send=S
        if cross is None:
            s=frombits((Es<<23)|(Send&0x7fffff)); base+=m
        else:
            k,Sb=cross
            sbf=frombits((Es<<23)|(Sb&0x7fffff))
            s=np.float32(sbf+t[base+k]); base+=k+1
    return s,passes


def exact_range(t, k0, k1, s, CH=64, E=2):
    """cells [k0, k1) added to s exactly (the warp's fallback): ordered_sum() started from s"""
    n = k1; base = k0
    while base < n:
        sb = fbits(s); Es = (sb >> 23) & 0xff
        if Es == 0 or Es == 255:
            s = np.float32(s + t[base]); base += 1; continue
        S = (sb & 0x7fffff) | 0x800000
        m = min(CH, n - base); crossed = False
        for k in range(m):
            q, gt, tie = elem(Es, fbits(t[base + k]))
            Sn = S + q + gt + (tie & ((S + q) & 1))
            if Sn >= 1 << 24:
                s = np.float32(frombits((Es << 23) | (S & 0x7fffff)) + t[base + k]); base += k + 1; crossed = True; break
            S = Sn
        if not crossed:
            s = frombits((Es << 23) | (S & 0x7fffff)); base += m
    return s


def stitched_sum(t, chunk=64):
    """returns (sum, number of chunks taken as one integer, number redone exactly)"""
    n = len(t)
    nch = (n + chunk - 1) // chunk
    partial = [np.float32(np.sum(t[c * chunk:(c + 1) * chunk], dtype=np.float64)) for c in range(nch)]   # any order will do
    info = []
    for c in range(1, nch):
        pre = np.float32(np.sum(partial[:c], dtype=np.float64))
        es = (fbits(pre) >> 23) & 0xff
        d = 0; tie_any = 0
        for x in t[c * chunk:(c + 1) * chunk]:
            q, gt, tie = elem(es, fbits(x))
            d = sat(d + q + gt); tie_any |= tie
        info.append((-1 if es in (0, 255) else es, -1 if tie_any else d))
    s = np.float32(0)
    for x in t[:chunk]:
        s = np.float32(s + x)
    fast = slow = 0
    for c in range(1, nch):
        es_c, d_c = info[c - 1]
        sb = fbits(s); S = (sb & 0x7fffff) | 0x800000
        if ((sb >> 23) & 0xff) == es_c and d_c >= 0 and S + d_c < 1 << 24:
            s = frombits((sb & 0x7f800000) | ((S + d_c) & 0x7fffff)); fast += 1
        else:
            s = exact_range(t, c * chunk, min(n, (c + 1) * chunk), s); slow += 1
    return s, fast, slow


def serial(t):
    s = np.float32(0)
    for x in t:
        s = np.float32(s + x)
    return s


def random_terms(kind, n, rng):
    """non-negative float32 terms of different characters: squares of normals, many exact ties, tiny + one huge, ..."""
    if kind == 0:
        return (rng.standard_normal(n) ** 2).astype(np.float32)
    if kind == 1:
        return (rng.integers(0, 64, n) / np.float32(8)).astype(np.float32)                     # many ties
    if kind == 2:
        return (rng.integers(0, 4, n) * np.float32(2.0 ** -int(rng.integers(0, 30)))).astype(np.float32)
    if kind == 3:
        return (10.0 ** rng.uniform(-12, 6, n)).astype(np.float32)
    t = np.where(rng.random(n) < 0.3, 0, rng.integers(1, 5, n) * np.float32(0.5)).astype(np.float32)
    t[rng.integers(0, n)] = np.float32(3e7)
    return t


def check(trials=300, seed=1, chunk=64, per_thread=4):
    rng = np.random.default_rng(seed)
    for trial in range(trials):
        n = int(rng.integers(1, 3000))
        t = random_terms(trial % 5, n, rng)
        a = serial(t)
        b, _ = ordered_sum(t, CH=chunk, E=per_thread)
        assert fbits(a) == fbits(b), (trial, trial % 5, n, a, b)
        c, _, _ = stitched_sum(t, chunk=int(rng.choice([16, 64, 256])))
        assert fbits(a) == fbits(c), ('stitched', trial, trial % 5, n, a, c)
    return trials


if __name__ == '__main__':
    print('model ok:', check(), 'trials')
