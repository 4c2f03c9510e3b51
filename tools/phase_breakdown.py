import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genefaceplusplus_b200 import _capi, scene as scn
from genefaceplusplus_b200.renderer import RADNeRFTorso
prec = sys.argv[1] if len(sys.argv)>1 else 'fp16'
if '--one-cta' in sys.argv: os.environ['GFPP_ONE_CTA'] = '1'   # phase timings without the co-resident CTA
sc = scn.Scene(H=512,W=512,T=20,torso=True,density_scale=8.0)
m = RADNeRFTorso(sc.hparams); m.load_state_dict(sc.state); m.density_scale=8.0; m.mlp_precision=prec; m=m.cuda().eval()
poses = torch.stack([sc.pose(t) for t in range(20)])
kw=dict(cond_seq=sc.cond,bg_color=sc.bg_color,bg_coords=sc.bg_coords,T_thresh=0.01,frames_per_call=20)
m.render_clip(poses, sc.intrinsics,512,512,**kw); torch.cuda.synchronize()
buf = torch.zeros(32,dtype=torch.int64,device='cuda')
_capi.lib().gfpp_profile_phases(buf.data_ptr())
m.render_clip(poses, sc.intrinsics,512,512,**kw); torch.cuda.synchronize()
_capi.lib().gfpp_profile_phases(None)
v = buf.cpu().tolist(); nb = v[31]
if not os.environ.get('GFPP_HEAD_V1') and prec in ('fp16', 'robust'):   # row-owner kernel: stamps of row 0 of slot 0 of every CTA
    names2 = ['operand rows (pos feats, cond, SH)', 'fate + march + refill', 'wait amb L0', 'epi amb L0', 'wait amb L1', 'epi amb L1', 'wait amb out',
              'tanh + ambient gather', 'prefetch pos levels 0-7', 'wait sig L0', 'epi sig L0', 'prefetch pos levels 8-15', 'wait sig L1', 'epi sig L1',
              'wait sig L2', 'epi sig L2 (sigma + geo)', 'wait col L0', 'epi col L0', 'wait col out', 'sigmoid + composite + batch check']
    tot = sum(v[:20])
    print(f'precision {prec} (v2): batches of slot 0 {nb}, cycles per batch of one slot = {tot/nb:.0f}')
    for n, c in zip(names2, v[:20]): print(f'  {n:28s} {c/nb:9.0f} cyc/batch  {100*c/tot:5.1f}%')
    sys.exit(0)
tot=sum(v[:11])
names=['refill+publish','pos gather+cond','MMA amb0','epi amb0','MMA amb1','epi amb1','narrow amb + tanh','amb gather','sigma net (3 MMA+3 epi)','color net (2 MMA + epi)','composite+march']
print(f'precision {prec}: batches {nb}, cycles/batch (thread 0 of each CTA) = {tot/nb:.0f}')
for n,c in zip(names,v[:11]): print(f'  {n:28s} {c/nb:9.0f} cyc/batch  {100*c/tot:5.1f}%')
