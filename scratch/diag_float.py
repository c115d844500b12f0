import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import CASES, load_case
from qfa_b200 import QFA
prec = sys.argv[1] if len(sys.argv)>1 else "fp32"
for name in CASES:
    c,g = load_case(name,"f64"); c32,g32 = load_case(name,"f32")
    Npix,Nh=c["F"].shape
    m=QFA(c["Nb"],Npix-c["Nb"],Nh,torch.device("cuda:0"),tau=c["law"],model_params={k:c[k] for k in ("F","Psi","omega","tau0","c0","beta")},precision=prec)
    m.mu=torch.tensor(c["mu"])
    d=lambda x: torch.as_tensor(x).cuda()
    o=m.predict_batch(d(c["flux"]),d(c["error"]),d(c["zabs"]),d(c["mask"]))
    npx=np.maximum(1,c["mask"].sum(1))
    e=np.abs(o["nll"].cpu().numpy()-g["pred_nll"]); eref=np.abs(g32["pred_nll"]-g["pred_nll"])
    loss,grads=m.forward(d(c["delta"]),d(c["error"]),d(c["zabs"]),d(c["mask"]))
    def rel(a,b):
        a=np.asarray(a,float);b=np.asarray(b,float);ok=~np.isnan(b); return np.abs(a[ok]-b[ok]).max()/max(np.abs(b[ok]).max(),1e-300)
    print(f"{name:8s} nll err max {e.max():.3e} (per px {np.max(e/npx):.2e}) ref32 err {np.nanmax(eref):.3e} | cont {rel(o['cont'].cpu().numpy(),g['pred_cont']):.1e} unc {rel(o['unc'].cpu().numpy(),g['pred_unc']):.1e} hm {rel(o['hmean'].cpu().numpy(),g['pred_hmean']):.1e} gF {rel(grads['F'].cpu().numpy(),g['grad_F']):.1e} ref32 gF {rel(g32['grad_F'],g['grad_F']):.1e} gPsi {rel(grads['Psi'].cpu().numpy(),g['grad_Psi']):.1e} gt0 {rel(grads['tau0'].cpu().numpy(),g['grad_tau0']):.1e}")
