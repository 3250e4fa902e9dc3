#!/usr/bin/env python
"""CPU emulation of split-precision tensor-core arithmetic for the policy/value network: how far are the priors from
an fp64 forward when every matmul is computed as hi.hi + hi.lo + lo.hi with operands split into TF32 / fp16 / bf16
halves?  ("Precision".)  Uses oracle/net_oracle.py for weights and the fp64 reference; no GPU needed."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import net_oracle as no
import torch.nn.functional as F
torch.set_num_threads(8)
params = no.random_params(0)
rng = np.random.default_rng(1)
B = 64
boards = np.zeros((B,81), np.int8); turns = np.zeros(B, np.int64)
for b in range(B):
    n = rng.integers(0, 60)
    cells = rng.permutation(81)[:n]
    for i,c in enumerate(cells): boards[b,c] = 1 + (i&1)
    turns[b] = n & 1
imgs = np.stack([no.encode_image(b,int(t)) for b,t in zip(boards,turns)])
P64, V64, L64 = no.forward(params, imgs, torch.float64)

def split_tf32(x):
    x32 = x.float()
    hi = (x32.view(torch.int32) & -8192).view(torch.float32)
    lo = x32 - hi
    lo = (lo.view(torch.int32) & -8192).view(torch.float32)  # tf32 truncation of lo by the MMA
    return hi.double(), lo.double()
def split_f16(x, scale=1.0):
    x32 = x.float()*scale
    hi = x32.half()
    lo = (x32 - hi.float()).half()
    return hi.double()/scale, lo.double()/scale
def split_bf16(x, scale=1.0):
    x32 = x.float()
    hi = x32.bfloat16()
    lo = (x32 - hi.float()).bfloat16()
    return hi.double(), lo.double()
def wscale(w):
    m = float(w.abs().max()); import math
    return 2.0**math.floor(math.log2(16384.0/m))
def mm(a, w, mode):
    # a [M,K] fp32 tensor, w [K,N]
    if mode == 'fp32': return (a.float() @ w.float()).double()
    if mode == 'tf32x3': ah, al = split_tf32(a); wh, wl = split_tf32(w)
    elif mode == 'f16x3': ah, al = split_f16(a); wh, wl = split_f16(w, wscale(w))
    elif mode == 'f16x3_noscale': ah, al = split_f16(a); wh, wl = split_f16(w)
    elif mode == 'bf16x3': ah, al = split_bf16(a); wh, wl = split_bf16(w)
    return ah@wh + ah@wl + al@wh
def fwd(mode):
    p = {name: torch.from_numpy(np.asarray(a)) for (name,_),a in zip(no.PARAM_SPECS, params)}
    x = torch.from_numpy(imgs).reshape(B,9,9,3)
    lr = lambda t: F.leaky_relu(t, 0.2)
    x = lr(x.reshape(-1,3) @ p['conv_w'][0,0] + p['conv_b']).float()   # [B*81,128]
    mx = [float(x.abs().max())]
    for i in range(3):
        h = lr((mm(x, p[f'res{i}_w0'][0,0], mode) + p[f'res{i}_b0']).float())
        hh = h.reshape(B,9,9,32).permute(0,3,1,2)
        dw = p[f'res{i}_dw']
        hh = F.conv2d(hh, dw[:,:,:,0].permute(2,0,1).unsqueeze(1), None, padding=1, groups=32)
        h = hh.permute(0,2,3,1).reshape(-1,32)
        mx.append(float(h.abs().max()))
        h = lr((mm(h, p[f'res{i}_pw'][0,0], mode) + p[f'res{i}_b1']).float())
        h = (mm(h, p[f'res{i}_w2'][0,0], mode) + p[f'res{i}_b2']).float()
        x = lr(h + x)
        mx.append(float(x.abs().max()))
    flat = x.reshape(B, 10368)
    h = lr((mm(flat, p['fc0_w'], mode) + p['fc0_b']).float())
    mx.append(float(h.abs().max()))
    h = lr((mm(h, p['fc1_w'], mode) + p['fc1_b']).float())
    mx.append(float(h.abs().max()))
    v = torch.tanh(h @ p['v_w'] + p['v_b']).reshape(B)
    logits = h @ p['p_w'] + p['p_b']
    return torch.softmax(logits.double(),1).numpy(), v.numpy(), logits.numpy(), mx
for mode in ['fp32','tf32x3','f16x3','f16x3_noscale','bf16x3']:
    P, V, L, mx = fwd(mode)
    big = P64 > 1e-6
    rel = np.abs(P-P64)[big]/P64[big]
    print(mode, 'prior rel max %.2e mean %.2e | logit abs max %.2e | v abs max %.2e' % (rel.max(), rel.mean(), np.abs(L-L64).max(), np.abs(V-V64).max()), 'act max', ['%.1f'%m for m in mx])
