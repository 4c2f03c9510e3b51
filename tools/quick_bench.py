import sys, os, torch, time
sys.path.insert(0, os.getcwd())
from genefaceplusplus_b200 import scene as scn
from genefaceplusplus_b200.renderer import RADNeRFTorso
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
sc = scn.Scene(H=512, W=512, T=50, torso=True, density_scale=8.0)
m = RADNeRFTorso(sc.hparams); m.load_state_dict(sc.state); m.density_scale = 8.0; m.mlp_precision = prec; m = m.cuda().eval()
poses = torch.stack([sc.pose(t) for t in range(50)]).cuda()
kw = dict(cond_seq=sc.cond.cuda(), bg_color=sc.bg_color.cuda(), bg_coords=sc.bg_coords.cuda(), T_thresh=0.01, frames_per_call=50)
out = torch.empty(50, 512*512, 3, device="cuda")
for _ in range(2): m.render_clip(poses, sc.intrinsics, 512, 512, out=out, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): m.render_clip(poses, sc.intrinsics, 512, 512, out=out, **kw)
e1.record(); torch.cuda.synchronize()
print(f"{prec} POLL={os.environ.get('GFPP_V2_POLL_NS')} STAG={os.environ.get('GFPP_V2_STAGGER')}: {150/(e0.elapsed_time(e1)/1000):.0f} fps")
