import torch, sys
sys.path.insert(0,'/root/repo')
import deep_staple_b200 as ds
from oracle import dp_loss as ol
dev=torch.device('cuda:0')
for C,shape,dtype in ((8,(4,8,8),torch.float32),(5,(3,5,7),torch.float16),(4,(16,24),torch.float32)):
    gen=torch.Generator().manual_seed(1000+C+len(shape))
    B,M=3,10
    logits=(2*torch.randn((B,C,*shape),generator=gen)).to(dtype)
    target=torch.randint(0,C,(B,*shape),generator=gen)
    emb_w=torch.randn((M,1),generator=gen); idx=torch.randperm(M,generator=gen)[:B]
    fixed=torch.rand(M,generator=gen)*5+1; cw=torch.rand(C,generator=gen)+0.5
    ref=ol.run_loss(logits,target,emb_w,idx,fixed,cw,True,8.0,True)
    for need_grad in (False, True):
        z=logits.to(dev).requires_grad_(need_grad)
        bare=emb_w.reshape(-1)[idx].to(dev).requires_grad_(True)
        out=ds.dp_weighted_loss(z,target.to(dev),bare,fixed[idx].to(dev),True,grad_scaler=8.0,class_weights=cw.to(dev),return_all=True)
        V=logits[0,0].numel()
        lf,gf,wf=ol.closed_form(out.inst_stats[:,:,0].sum(1).cpu()/V, out.inst_stats[:,1:,2].sum(1).cpu(), emb_w.reshape(-1)[idx], fixed[idx], V, True, 8.0)
        print(C,shape,dtype,'grad' if need_grad else 'nograd','loss',out.dp_loss.item(),'ref',ref.dp_loss.item(),'closed-form-from-our-stats',lf.item())
        print('  per_inst_ce',out.per_instance_ce.cpu().tolist(), ref.per_inst_ce.tolist())
        print('  fg',out.fg_pred.cpu().tolist(), ref.fg_pred.tolist(),'w',out.instance_weight.cpu().tolist(), wf.tolist())
