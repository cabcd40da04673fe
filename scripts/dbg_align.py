import torch, os, sys
sys.path.insert(0,'/root/repo')
sys.path.insert(0,'/root/repo/tests'); from test_align_gpu import fx, run_slam, cpu
dev=torch.device('cuda:0')
for name in ["align_match3.pt","align_dust3r3.pt"]:
    f=fx(name)
    for tag,(n1,n2) in {"short":f["niter"],"full":(300,200)}.items():
        _,rc_,rf_,_=run_slam(f,dev,n1,n2)
        for which,res,ref in (("coarse",cpu(rc_),f["out"][tag]["coarse"]),("fine",cpu(rf_),f["out"][tag]["fine"])):
            dK=(res["intrinsics"]-ref["intrinsics"]).abs().max().item()
            dd=max(((a.ravel()-b.ravel()).abs()/b.ravel().abs()).max().item() for a,b in zip(res["depthmaps"],ref["depthmaps"]))
            rel=torch.linalg.inv(res["cam2w"][0:1])@res["cam2w"]; rrel=torch.linalg.inv(ref["cam2w"][0:1])@ref["cam2w"]
            dr=(rel-rrel).abs().max().item()
            G=ref["cam2w"][0]@torch.linalg.inv(res["cam2w"][0])
            dp=max(((a@G[:3,:3].T+G[:3,3])-b).abs().max().item() for a,b in zip(res["pts3d"],ref["pts3d"]))
            print(name,tag,which,"dK %.2e ddepth_rel %.2e drelpose %.2e dpts %.2e"%(dK,dd,dr,dp))
