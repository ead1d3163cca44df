"""Numerics of a tensor-core 1024-point DFT (two 32-point stages as [64 x 64] real GEMMs, operands split into FP16 or BF16 hi + lo,
FP32 accumulation) against a float64 FFT: the EVM figures quoted in DESIGN.md 5.2.3.  Pure numpy; runs anywhere."""
import numpy as np
rng=np.random.default_rng(1)
F=8; N=1024
x=(rng.standard_normal((F,N))+1j*rng.standard_normal((F,N))).astype(np.complex64)
def split16(a):
    hi=a.astype(np.float16); lo=(a-hi.astype(np.float32)).astype(np.float16); return hi,lo
def splitbf(a):
    u=a.view(np.uint32)&np.uint32(0xffff0000); hi=u.view(np.float32); lo=((a-hi).view(np.uint32)&np.uint32(0xffff0000)).view(np.float32); return hi,lo
def gemm3(A,B,split,terms=3):
    # A: data (rows x K) fp32, B: matrix K x N fp32 ; fp32 accumulate
    Ah,Al=split(A); Bh,Bl=split(B)
    f=lambda u,v: (u.astype(np.float32)@v.astype(np.float32)).astype(np.float32)
    r=f(Ah,Bh)
    if terms>=2: r=r+f(Al,Bh)
    if terms>=3: r=r+f(Ah,Bl)
    return r
def dftmat(R):
    j=np.arange(R); W=np.exp(-2j*np.pi*np.outer(j,j)/R)
    M=np.block([[W.real, W.imag],[-W.imag, W.real]])   # [xr xi] @ M = [yr yi]  with y = x@W
    return M.astype(np.float32)
M32=dftmat(32)
def fft_tc(x,split,terms=3,scale_pow=0):
    out=np.empty_like(x)
    for f in range(x.shape[0]):
        X=x[f].reshape(32,32)            # X[n1][n2]
        A=np.concatenate([X.T.real,X.T.imag],axis=1).astype(np.float32)*np.float32(2.0**scale_pow)   # rows n2, K=(n1 re | n1 im)
        Y=gemm3(A,M32,split,terms)       # rows n2, cols (k1 re | k1 im)
        Yc=(Y[:,:32]+1j*Y[:,32:]).astype(np.complex64)   # [n2][k1]
        tw=np.exp(-2j*np.pi*np.outer(np.arange(32),np.arange(32))/1024).astype(np.complex64)
        Yc=(Yc*tw).astype(np.complex64)
        A2=np.concatenate([Yc.T.real,Yc.T.imag],axis=1).astype(np.float32)  # rows k1, K=n2
        Z=gemm3(A2,M32,split,terms)      # rows k1, cols k2
        Zc=(Z[:,:32]+1j*Z[:,32:])        # [k1][k2] -> k=k1+32k2
        out[f]=(Zc.T.reshape(-1)*np.float32(2.0**-scale_pow)).astype(np.complex64)
    return out
ref=np.fft.fft(x.astype(np.complex128),axis=1)
def evm(a): return np.sqrt(np.sum(np.abs(a-ref)**2)/np.sum(np.abs(ref)**2)), np.max(np.abs(a-ref))/np.sqrt(np.mean(np.abs(ref)**2))
print("numpy c64 fft ",evm(np.fft.fft(x,axis=1).astype(np.complex64)))
import scipy.fft
print("scipy c64 fft ",evm(scipy.fft.fft(x,axis=1)))
for sp in (0,-4,4):
  print("fp16x3 scale 2^%d"%sp,evm(fft_tc(x,split16,3,sp)))
print("fp16x2",evm(fft_tc(x,split16,2)))
print("fp16x1",evm(fft_tc(x,split16,1)))
print("bf16x3",evm(fft_tc(x,splitbf,3)))
