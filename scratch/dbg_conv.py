import numpy as np, torch, sys
sys.path.insert(0, '.')
from synthobs_b200 import _cabi
from synthobs_b200.morph import images
from oracle import image_oracle
rng = np.random.default_rng(2)
a = rng.random((120, 120)); k = rng.random((33, 33))
ref = image_oracle.convolve_fft_same(a, k)
x = torch.from_numpy(a).cuda().float().reshape(1,120,120)
kk = torch.from_numpy(k/k.sum()).cuda().float()
import scipy.fft as sfft
P = sfft.next_fast_len(120+33-1, real=True)
A = torch.fft.rfft2(x, s=(P,P)); K = torch.fft.rfft2(kk, s=(P,P))
print('A', A.shape, A.is_contiguous(), A.stride(), 'K', K.shape, K.is_contiguous(), K.stride())
full = torch.fft.irfft2(A*K, s=(P,P))
got = full[0, 16:16+120, 16:16+120].double().cpu().numpy()
print('torch mul err', np.abs(got-ref).max()/np.abs(ref).max())
prod = torch.empty_like(A)
rc = _cabi.lib.so_complex_mul_bcast(_cabi.ptr(torch.view_as_real(A)), _cabi.ptr(torch.view_as_real(K)), _cabi.ptr(torch.view_as_real(prod)), A.shape[-2]*A.shape[-1], A.shape[0], _cabi.stream())
torch.cuda.synchronize()
print('rc', rc, 'kernel vs torch', (prod - A*K).abs().max().item(), (A*K).abs().max().item())
got2 = images.convolve_fft(a, k)
print('api err', np.abs(got2-ref).max()/np.abs(ref).max())
