import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import domain_specific_image_compression_b200 as sic
from oracle import torch_port as TP, numpy_ref as R
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
G=np.load('tests/golden/model_small.npz')
m=sic.CompressionModel(N=16,M=24,min_nu=2.0).cuda()
m.load_state_dict({k[3:]:torch.from_numpy(G[k]) for k in G.files if k.startswith('sd.')})
m.eval(); x=torch.from_numpy(G['x']).cuda()
with torch.no_grad():
    out=m(x,'round'); loss,Rr,D=sic.rate_distortion_loss(out,x,100.0,'mse')
    sd={k:v.detach() for k,v in m.state_dict().items()}
    ref=TP.forward(sd,x,'round',training=False)
for k in ('y','z','y_tilde','z_tilde','sigma','nu','nll_y','nll_z','x_hat'):
    a=out[k].cpu().numpy(); b=ref[k].cpu().numpy(); g=G['eval.'+k]
    print(k, "ours-vs-port max", np.abs(a-b).max(), "ours-vs-golden max", np.abs(a-g).max(), "sum ours", a.sum(), "port", b.sum(), "golden", g.sum())
print("bits_y", out['nll_y']._sic_bits, "bits_z", out['nll_z']._sic_bits, "R", float(Rr), "golden R", G['eval.R'])
print("nll_y oracle on gpu inputs:", R.studentt_nll_f64(out['y_tilde'].cpu().numpy(), out['sigma'].cpu().numpy(), out['nu'].cpu().numpy()).sum())
