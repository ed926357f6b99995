"""Index-math model of the multi-pass NTT kernel (small prime field for speed)."""
import random
P = 2013265921  # 15*2^27+1
G = 31
def root(n): return pow(G, (P-1)//n, P)
def dft(a, w):
    n=len(a); return [sum(a[j]*pow(w,i*j,P) for j in range(n))%P for i in range(n)]

def plan(k, maxb=9):
    # split k into passes of <= maxb bits, as even as possible
    npass = -(-k//maxb)
    base, rem = divmod(k, npass)
    return [base+1]*rem + [base]*(npass-rem)

def rounds(b):
    # split b into rounds of <=3 bits, big first
    r=[]
    while b>0:
        t=min(3,b); r.append(t); b-=t
    return r

def small_dft_inreg(v, w8tab_root):
    # DFT of len(v)=2^r with root w
    return dft(v, w8tab_root)

def ntt_pass(src, k, omega, bits, p):
    """pass p (0-based) : [j_p][J][I] -> [J][i_p][I] with twiddle."""
    n=1<<k
    b=bits[p]; npp=1<<b
    I=1<<sum(bits[:p]); J=n//(npp*I)
    dst=[None]*n
    w_np = pow(omega, n//npp, P)           # root of the size-n_p transform
    rs=rounds(b)
    for col in range(J*I):               # flat column m = Jv*I + Iv
        Jv, Iv = divmod(col, I)
        # gather the column
        x=[src[jp*(n//npp)+col] for jp in range(npp)]
        # in-tile rounds, in place on position array
        # position digits (x1,x2,x3), x1 most significant
        sizes=[1<<r for r in rs]
        nr=len(rs)
        # strides
        strides=[1]*nr
        for q in range(nr-2,-1,-1): strides[q]=strides[q+1]*sizes[q+1]
        for q in range(nr):
            sq=sizes[q]; st=strides[q]
            wq=pow(w_np, npp//sq, P)   # root of size sq
            rest = npp//sq
            for other in range(rest):
                # decompose 'other' into digits excluding q -> base position
                # digits higher than q: hi, lower than q: lo
                lo = other % st
                hi = other // st
                basepos = hi*st*sq + lo
                v=[x[basepos+a*st] for a in range(sq)]
                V=dft(v,wq)
                # twiddle: w_{sq*st}^{c_q * lo}  where lo = remaining lower digits value (a_{q+1..})
                wt = pow(w_np, npp//(sq*st), P)
                for c in range(sq):
                    x[basepos+c*st]=V[c]*pow(wt,c*lo,P)%P
        # now x[pos] with pos=(c1,c2,c3) holds output index i = c1 + s1*c2 + s1*s2*c3
        for pos in range(npp):
            rem=pos; ip=0; mult=1
            digs=[]
            for q in range(nr):
                d = rem//strides[q]; rem%=strides[q]; digs.append(d)
            for q in range(nr):
                ip += digs[q]*mult; mult*=sizes[q]
            # inter-pass twiddle (omega^I)^(ip*Jv)
            val = x[pos]*pow(omega, I*ip*Jv, P)%P
            dst[Jv*npp*I + ip*I + Iv]=val
    return dst

def ntt(a,k,omega,maxb=9):
    bits=plan(k,maxb)
    cur=a
    for p in range(len(bits)):
        cur=ntt_pass(cur,k,omega,bits,p)
    return cur

random.seed(1)
for k,maxb in [(3,9),(4,2),(5,3),(6,3),(7,4),(8,4),(9,9),(10,5),(10,4),(9,3),(11,4)]:
    n=1<<k; w=root(n)
    a=[random.randrange(P) for _ in range(n)]
    assert ntt(a,k,w,maxb)==dft(a,w),(k,maxb)
    print(k,maxb,plan(k,maxb),'ok')
