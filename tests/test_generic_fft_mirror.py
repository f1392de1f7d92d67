"""CPU mirror of the generic kernel's three-stages-per-pass FFT (fhe_string_bounty_b200/csrc/pbs_generic.cu, struct Fft8): the same pass
structure restated in numpy -- leftover radix-2 / radix-4 pass first, then passes q = M'/8, M'/64, ..., 1 with butterfly b -> (j, i0),
the constants W_8^r / W_4^r between the stages and the stage twiddles applied once at the end as w1^brev3(a) -- checked against
numpy's FFT (forward: natural in, bit-reversed out; inverse: unscaled round trip), plus the slot swizzle: a bijection that spreads
every quarter-warp access of every pass over all eight 16-byte bank groups.  No GPU needed; the GPU parity tests
(tests/test_gpu_param_sets.py) run the real kernel."""
import numpy as np
import pytest

S=2**-0.5
def W(n,e): return np.exp(-2j*np.pi*e/n)
def brev(x,bits): return int(format(x,'0%db'%bits)[::-1],2)
def fwd(buf):
    M=len(buf); L=M.bit_length()-1; R=L%3; buf=buf.copy()
    if R==1:
        for b in range(M//2):
            u,v=buf[b],buf[b+M//2]; buf[b]=u+v; buf[b+M//2]=(u-v)*W(M,b)
    elif R==2:
        q=M//4
        for b in range(M//4):
            a0,a1,a2,a3=buf[b],buf[b+q],buf[b+2*q],buf[b+3*q]
            b0,b1,b2,b3=a0+a2,a1+a3,a0-a2,-1j*(a1-a3)
            w1=W(M,b); w2=w1*w1
            buf[b]=b0+b1; buf[b+q]=(b0-b1)*w2; buf[b+2*q]=(b2+b3)*w1; buf[b+3*q]=(b2-b3)*w1*w2
    MP=M>>R; q=MP//8
    while q>=1:
        for b in range(M//8):
            j=b&(q-1); i0=((b-j)<<3)+j
            x=[buf[i0+m*q] for m in range(8)]
            for a in range(4):
                u=x[a]; d=x[a]-x[a+4]; x[a]=u+x[a+4]
                if a==0: r=d
                elif a==1: r=complex((d.real+d.imag)*S,(d.imag-d.real)*S)
                elif a==2: r=-1j*d
                else: r=complex((d.imag-d.real)*S,-(d.real+d.imag)*S)
                x[a+4]=r
            for base in (0,4):
                u0,u1=x[base],x[base+1]
                d0=u0-x[base+2]; d1=-1j*(u1-x[base+3])
                x[base]=u0+x[base+2]; x[base+1]=u1+x[base+3]; x[base+2]=d0; x[base+3]=d1
            for base in (0,2,4,6):
                u,v=x[base],x[base+1]; x[base]=u+v; x[base+1]=u-v
            if q>1:
                w1=W(8*q,j); w2=w1*w1; w3=w1*w2; w4=w2*w2
                x[4]*=w1; x[2]*=w2; x[6]*=w3; x[1]*=w4; x[5]*=w4*w1; x[3]*=w4*w2; x[7]*=w4*w3
            for m in range(8): buf[i0+m*q]=x[m]
        q//=8
    return buf
def inv(buf):
    M=len(buf); L=M.bit_length()-1; R=L%3; buf=buf.copy()
    MP=M>>R; QMAX=MP//8; q=1
    while q<=QMAX:
        for b in range(M//8):
            j=b&(q-1); i0=((b-j)<<3)+j
            x=[buf[i0+m*q] for m in range(8)]
            if q>1:
                w1=W(8*q,j); w2=w1*w1; w3=w1*w2; w4=w2*w2
                c=np.conj
                x[4]*=c(w1); x[2]*=c(w2); x[6]*=c(w3); x[1]*=c(w4); x[5]*=c(w4*w1); x[3]*=c(w4*w2); x[7]*=c(w4*w3)
            for base in (0,2,4,6):
                u,v=x[base],x[base+1]; x[base]=u+v; x[base+1]=u-v
            for base in (0,4):
                u0,u1,v0,v1=x[base],x[base+1],x[base+2],1j*x[base+3]
                x[base]=u0+v0; x[base+2]=u0-v0; x[base+1]=u1+v1; x[base+3]=u1-v1
            for a in range(4):
                u=x[a]; d=x[a+4]
                if a==0: v=d
                elif a==1: v=complex((d.real-d.imag)*S,(d.real+d.imag)*S)
                elif a==2: v=1j*d
                else: v=complex(-(d.real+d.imag)*S,(d.real-d.imag)*S)
                x[a]=u+v; x[a+4]=u-v
            for m in range(8): buf[i0+m*q]=x[m]
        q*=8
    if R==1:
        for b in range(M//2):
            u=buf[b]; v=buf[b+M//2]*np.conj(W(M,b)); buf[b]=u+v; buf[b+M//2]=u-v
    elif R==2:
        q=M//4
        for b in range(M//4):
            w1=W(M,b); w2=w1*w1; c=np.conj
            c0,c1,c2,c3=buf[b],buf[b+q]*c(w2),buf[b+2*q]*c(w1),buf[b+3*q]*c(w1)*c(w2)
            b0,b1,b2,b3=c0+c1,c0-c1,c2+c3,1j*(c2-c3)
            buf[b]=b0+b2; buf[b+2*q]=b0-b2; buf[b+q]=b1+b3; buf[b+3*q]=b1-b3
    return buf


@pytest.mark.parametrize("M", [64, 128, 256, 512, 1024, 4096])
def test_fft8_pass_structure_matches_numpy(M):
    rng = np.random.default_rng(M)
    x = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    L = M.bit_length() - 1
    y = fwd(x)
    ref = np.fft.fft(x)
    want = np.array([ref[brev(p, L)] for p in range(M)])
    assert np.abs(y - want).max() < 1e-10 * np.sqrt(M)
    assert np.abs(inv(y) / M - x).max() < 1e-12


def test_fft8_slot_swizzle_is_conflict_free():
    sw = lambda i: i ^ ((i >> 3) & 7)
    assert sorted(sw(i) for i in range(4096)) == list(range(4096))
    for q in (1, 8, 64, 512):
        for m in range(8):                         # one LDS.128 / STS.128 of the unrolled butterfly
            for b0 in range(0, 512, 8):            # a quarter-warp = 8 consecutive butterflies
                groups = set()
                for b in range(b0, b0 + 8):
                    j = b & (q - 1)
                    i0 = ((b - j) << 3) + j
                    groups.add(sw(i0 + m * q) & 7)
                assert len(groups) == 8, (q, m, b0)
    # consecutive elements (fold / multiply-accumulate / unfold accesses)
    for base in range(0, 4096, 8):
        assert len({sw(base + e) & 7 for e in range(8)}) == 8


def test_fft4_slot_swizzle_is_conflict_free():
    """the two-stages-per-pass transform (k >= 2 parameter sets): slot = i ^ ((i >> 3) & 3) ^ ((i >> 2) & 4)"""
    sw = lambda i: i ^ ((i >> 3) & 3) ^ ((i >> 2) & 4)
    assert sorted(sw(i) for i in range(2048)) == list(range(2048))
    for q in (1, 4, 16, 64, 256):
        for m in range(4):
            for b0 in range(0, 512, 8):
                groups = set()
                for b in range(b0, b0 + 8):
                    j = b & (q - 1)
                    i0 = ((b - j) << 2) + j
                    groups.add(sw(i0 + m * q) & 7)
                assert len(groups) == 8, (q, m, b0)
