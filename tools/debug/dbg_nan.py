import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch
import durf_test_helpers as H, ref_cases as C
from oracle import durf_oracle as O
from durf_b200 import ops, _lib as L
from durf_b200.obbpose_model import MipNerfModel
from test_ref_golden import load, T
g = load('model')
name='c4_k8_overlap'
skw, mover, akw = C.MODEL_CASES[name]
sc = C._scene(**skw)
keep=[]
cfg = O.ModelConfig()
ret = O.model_forward(H.oracle_params(sc), H.oracle_rays(sc), T(sc['ext']), akw['ts'], False, False, False, 10.0, cfg=cfg, keep_raw=keep)
raw_rgb, raw_den, enc = keep[0]
print('oracle L0: enc nan', int(torch.isnan(enc).sum()), 'inf', int(torch.isinf(enc).sum()), 'raw_rgb nan', int(torch.isnan(raw_rgb).sum()), 'raw_den nan', int(torch.isnan(raw_den).sum()), 'comp nan', int(torch.isnan(ret[0].comp_rgb).sum()), 'w nan', int(torch.isnan(ret[0].weights).sum()), 'acc nan', int(torch.isnan(ret[0].acc).sum()))
nh = ret[0].dyn_mask.reshape(-1)
print('nhit>=2 rays', int((nh>=2).sum()), 'nan rays', int(torch.isnan(ret[0].comp_rgb).any(-1).sum()))
bad = torch.isnan(ret[0].comp_rgb).any(-1)
print('nan rays subset of multi-hit:', bool((bad <= (nh>=2)).all()), 'multi-hit w/o nan', int(((nh>=2)&~bad).sum()))
i = int(torch.nonzero(bad)[0])
print('ray', i, 'enc row nan/inf', int(torch.isnan(enc[i]).sum()), int(torch.isinf(enc[i]).sum()), 'enc max', float(enc[i][torch.isfinite(enc[i])].abs().max()))
# cuda
model = MipNerfModel(precision='fp32', num_objects=8)
v = H.cuda_variables(sc, model)
r = H.cuda_rays(sc)
box = v.box_centers[akw['ts']].contiguous()
fe = ops.obb_frontend(r.origins, r.directions, box, torch.from_numpy(sc['ext']).cuda())
bg_mult = 1.0 - fe['nhit']
rm = ops.raymarch(fe['origins_s'], fe['dirs_s'], r.radii.reshape(-1), 128, near=r.near, far=r.far, contract=True, ray_mult=bg_mult, want_gaussians=True)
f = rm['features'].cpu()
print('cuda enc nan', int(torch.isnan(f).sum()), 'inf', int(torch.isinf(f).sum()), 'row', int(torch.isnan(f[i]).sum()), int(torch.isinf(f[i]).sum()), 'max', float(f[i][torch.isfinite(f[i])].abs().max()))
print('cov diag min (cuda)', float(rm['cov_diag'][i].min()), 'bg_mult', float(bg_mult[i]))
d = (f[i]-enc[i]); print('enc diff finite max', float(d[torch.isfinite(d)].abs().max()) if torch.isfinite(d).any() else None)
viewenc = ops.viewdir_enc(r.viewdirs, 4)
rgb, den, _ = ops.mlp_fwd(model.bg_topology(), rm['features'], viewenc, v.blob('MLP_0'), M=sc['B'], N=128, precision=L.PREC_FP32)
print('cuda raw_rgb nan', int(torch.isnan(rgb).sum()), 'raw_den nan', int(torch.isnan(den).sum()), 'inf', int(torch.isinf(rgb).sum()))
got = model.apply(v, None, r, None, torch.from_numpy(sc['ext']).cuda(), torch.tensor([akw['ts']]), False, False, False, 10.0)
print('cuda comp nan', int(torch.isnan(got[0][0]).sum()), 'w nan', int(torch.isnan(got[0][3]).sum()))
print('ref comp nan', int(np.isnan(g[name+'/L0/comp_rgb']).sum()), 'ref weights nan', int(np.isnan(g[name+'/L0/weights']).sum()), 'ref acc nan', int(np.isnan(g[name+'/L0/acc']).sum()), 'dist nan', int(np.isnan(g[name+'/L0/distance']).sum()))
