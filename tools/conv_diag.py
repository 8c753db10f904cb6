import sys, torch
sys.path.insert(0, "/root/repo")
from speech_b200 import ops
def _build(specs, in_c=1):
    layers = []
    for out_c, h, w, s in specs:
        layers += [torch.nn.Conv2d(in_c, out_c, (h, w), stride=(s, s)), torch.nn.ReLU()]
        in_c = out_c
    return torch.nn.Sequential(*layers)
for specs,B,T,F in [([[32,5,32,2]],4,100,40), ([[8,5,8,2],[8,5,8,2]],3,61,80)]:
    torch.manual_seed(0)
    conv=_build(specs); x=torch.randn(B,T,F)
    c64=_build(specs).double(); c64.load_state_dict({k:v.double() for k,v in conv.state_dict().items()})
    y64=c64(x.double().unsqueeze(1)); b,c,t,f=y64.shape; y64=y64.transpose(1,2).reshape(b,t,c*f)
    w=torch.randn_like(y64); (y64*w).sum().backward()
    cc=conv.cuda(); y=ops.conv_stack(x.cuda(),cc,True); (y*w.float().cuda()).sum().backward()
    print("fwd err", (y.double().cpu()-y64).abs().max().item(), "scale", y64.abs().max().item())
    for (n,p64),(_,pc) in zip(c64.named_parameters(), cc.named_parameters()):
        ref=p64.grad; g=pc.grad.double().cpu()
        print(n, tuple(ref.shape), "err %.4g refmax %.4g  ratio(g/ref) median %.3f" % ((g-ref).abs().max().item(), ref.abs().max().item(), (g/ref).flatten().median().item()))
