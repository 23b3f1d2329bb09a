#!/usr/bin/env python
"""CPU feasibility study behind the packed-half GELU epilogue of the FFN-up GEMM (ce_gemm.cu: gelu_erf_h2).

Emulates the Abramowitz-Stegun erf-GELU in float16 arithmetic (NumPy rounds every operation to float16) and measures
(1) the per-activation error after the fp16 store against the exact erf-GELU and (2) the effect on the final sigmoid
scores of the MiniLM-L6-shaped scorer when ONLY the activation is replaced (everything else fp64, oracle/cross_encoder.py).
Result recorded in DESIGN.md K5:  fp32 formula 6.6e-9, float16 A-S formula 5.7e-5, float16 tanh-fit formula (round 2, the
one in the kernel) see the last line printed -- relative on the scores (tolerance 1e-3).
Uses oracle/ -- test infrastructure, not part of the product."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cross_encoder as ce
from sentio_b200.cross_encoder import MINILM_L6, CrossEncoderWeights
from sentio_b200.index import hash_tokenize_pairs
from sentio_b200 import synth

def gelu_as_f32(x):
    x=x.astype(np.float32)
    z=np.abs(x)*np.float32(0.70710678)
    t=np.float32(1)/(np.float32(1)+np.float32(0.3275911)*z)
    p=np.float32(1.061405429)*t+np.float32(-1.453152027)
    p=p*t+np.float32(1.421413741); p=p*t+np.float32(-0.284496736); p=p*t+np.float32(0.254829592)
    e=np.float32(1)-p*t*np.exp(-z*z)
    return (np.float32(0.5)*x*(np.float32(1)+np.copysign(e,x))).astype(np.float64)

def gelu_as_f16(x):
    h=np.float16
    x16=x.astype(np.float32).astype(h)           # bias add in fp32, then round to fp16
    z=(np.abs(x16)*h(0.70710678)).astype(h)
    t=(h(1)/((h(1)+(h(0.3275911)*z).astype(h)).astype(h))).astype(h)
    p=((h(1.061405429)*t).astype(h)+h(-1.453152027)).astype(h)
    for c in (1.421413741,-0.284496736,0.254829592):
        p=((p*t).astype(h)+h(c)).astype(h)
    ex=np.exp(-(z*z).astype(h).astype(np.float32)).astype(h)
    e=(h(1)-((p*t).astype(h)*ex).astype(h)).astype(h)
    return ((h(0.5)*x16).astype(h)*(h(1)+np.copysign(e,x16)).astype(h)).astype(h).astype(np.float64)

def gelu_tanhfit_f16(x):
    """the form the kernel uses since round 2 (scripts/fit_gelu.py): 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))), packed half"""
    h=np.float16
    x16=x.astype(np.float32).astype(h)
    x2=np.minimum((x16*x16).astype(h),h(36.0))
    q=((h(-0.00035873236644)*x2).astype(h)+h(0.0370503451315)).astype(h)
    q=((q.astype(np.float32)*x2.astype(np.float32))+np.float32(h(0.79745847075))).astype(h)
    u=(x16*q).astype(h)
    t=np.tanh(u.astype(np.float64))
    t=(t*(1+np.random.default_rng(0).uniform(-2.0**-11,2.0**-11,size=t.shape))).astype(h)   # tanh.approx.f16x2 error bound
    hx=(h(0.5)*x16).astype(h)
    return ((hx.astype(np.float32)*t.astype(np.float32))+hx.astype(np.float32)).astype(h).astype(np.float64)

xs=np.linspace(-6,6,200001)
exact=0.5*xs*(1+np.vectorize(math.erf)(xs/math.sqrt(2)))
VARIANTS=[("f32 A&S",gelu_as_f32),("f16 A&S",gelu_as_f16),("f16 tanh-fit",gelu_tanhfit_f16)]
for name,f in VARIANTS:
    y=f(xs); y16=y.astype(np.float16).astype(np.float64)
    err=np.abs(y16-exact)
    rel=err/np.maximum(np.abs(exact),1e-3)
    print(name,"max abs err after fp16 store",err.max(),"max rel (|g|>1e-3)",rel.max(),"rms abs",np.sqrt((err**2).mean()))
# end-to-end effect on the scorer (fp64 everything else)
w=CrossEncoderWeights.random(MINILM_L6,seed=0)
flat,off=synth.text_corpus_tokens(24,vocab=4000)
docs=synth.texts_from_tokens(flat,off)
ids,tt,lens=hash_tokenize_pairs("w1 w5 w9 w100 w3 w7",docs,128)
base=ce.numpy_forward(w,ids,tt,lens)[1]
for name,f in VARIANTS:
    old=ce._gelu; ce._gelu=lambda x,f=f: f(x)
    got=ce.numpy_forward(w,ids,tt,lens)[1]; ce._gelu=old
    print(name,"max rel err of sigmoid scores vs exact-erf forward:",np.max(np.abs(got-base)/base))
