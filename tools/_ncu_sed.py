import sys; sys.path.insert(0, '.')
import numpy as np, torch
import bench
from synthobs_b200 import synthetic
from synthobs_b200.sed import models
dev = torch.device('cuda', 0)
model, F, cosmo = bench.setup_model()
b = synthetic.galaxy_batch(256, fixed_n=100000, seed=100)
keys = ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')
d = {k: torch.from_numpy(b[k]).to(dev) for k in keys}; off = torch.from_numpy(b['offsets']).to(dev)
for _ in range(2):
    models.generate_Fnu_batch(model, F, d['Masses'], d['Ages'], d['Metallicities'], off, tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'])
torch.cuda.synchronize()
