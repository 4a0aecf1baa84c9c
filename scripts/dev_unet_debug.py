"""Dev (GPU): where does the UNet error come from? mine(bf16 kernels) vs fp32 oracle vs torch-bf16 oracle on GPU."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import faithdiff_dev_b200 as fd
from oracle import faithdiff_oracle as O
BF = torch.bfloat16
def rel(a, b): return ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()
torch.manual_seed(0)
cfg = O.TINY_UNET_CONFIG
ora = O.build_unet(cfg, lq_channels=64, encoder_kwargs=O.TINY_ENCODER_CONFIG)
for p in ora.parameters(): p.data = p.data.to(BF).float()
unet = fd.UNet2DConditionModel(**cfg); unet.load_state_dict(ora.state_dict())
ora_bf = copy.deepcopy(ora).cuda().to(BF)
for (h, w, t) in [(16, 16, 951.0), (16, 16, 1.0), (32, 32, 951.0), (32, 32, 1.0), (18, 14, 501.0), (64, 64, 501.0)]:
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, h, w, generator=g).to(BF).float(); lq = torch.randn(2, 64, h, w, generator=g).to(BF).float()
    ehs = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g).to(BF).float(); te = torch.randn(2, 64, generator=g).to(BF).float()
    ids = torch.tensor([[128., 128., 0., 8., 128., 128.]] * 2)
    outs = {}
    acts = {}
    def mk(name):
        def hook(m, i, o):
            acts[name] = (o[0] if isinstance(o, tuple) else o).detach()
        return hook
    hs = [ora.down_blocks[0].register_forward_hook(mk("down0")), ora.down_blocks[1].register_forward_hook(mk("down1")),
          ora.down_blocks[2].register_forward_hook(mk("down2")), ora.mid_block.register_forward_hook(mk("mid")),
          ora.up_blocks[0].register_forward_hook(mk("up0")), ora.up_blocks[1].register_forward_hook(mk("up1")),
          ora.up_blocks[2].register_forward_hook(mk("up2")), ora.conv_in.register_forward_hook(mk("conv_in")),
          ora.spatial_ch_projs.register_forward_hook(mk("proj")), ora.condition_embedding.register_forward_hook(mk("cond")),
          ora.down_blocks[0].resnets[0].register_forward_hook(mk("d0r0")),
          ora.down_blocks[1].resnets[0].register_forward_hook(mk("d1r0")), ora.down_blocks[1].attentions[0].register_forward_hook(mk("d1a0"))]
    with torch.no_grad():
        ref = ora(x, torch.tensor(t), encoder_hidden_states=ehs, input_embedding=lq, added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)[0]
        rbf = ora_bf(x.cuda().to(BF), torch.tensor(t), encoder_hidden_states=ehs.cuda().to(BF), input_embedding=lq.cuda().to(BF),
                     added_cond_kwargs={"text_embeds": te.cuda().to(BF), "time_ids": ids.cuda().to(BF)}, return_dict=False)[0]
    for hh in hs: hh.remove()
    out = unet(x.cuda().to(BF), torch.tensor(t), encoder_hidden_states=ehs.cuda().to(BF), input_embedding=lq.cuda().to(BF),
               added_cond_kwargs={"text_embeds": te.cuda().to(BF), "time_ids": ids.cuda()}, return_dict=False)[0]
    print("h=%d w=%d t=%g : mine %.3e | torch-bf16 %.3e" % (h, w, t, rel(out, ref), rel(rbf, ref)), flush=True)
    # stage-wise for mine
    from faithdiff_dev_b200 import ops
    temb = unet.time_embeddings(torch.tensor(t), te.cuda().to(BF), ids.cuda(), 2)
    kv, lk = unet.text_kv(ehs.cuda().to(BF))
    cond = unet.condition_embed(ops.nchw_to_nhwc(lq.cuda().to(BF)))
    print("   cond  %.3e" % rel(cond.permute(0, 3, 1, 2), acts["cond"]))
    xs = unet.stem(x.cuda().to(BF), cond)
    stem_ref = acts["conv_in"] + acts["proj"].transpose(1, 2).reshape(acts["conv_in"].shape)
    print("   stem  %.3e" % rel(xs.permute(0, 3, 1, 2), stem_ref))
    r0 = unet.down[0]["resnets"][0](xs, temb)
    print("   d0r0  %.3e" % rel(r0.permute(0, 3, 1, 2), acts["d0r0"]))
    # run from oracle's exact inputs through single blocks
    xin = ops.nchw_to_nhwc(acts["down0"].cuda().to(BF))
    r = unet.down[1]["resnets"][0](xin, temb)
    print("   d1r0 (oracle input) %.3e" % rel(r.permute(0, 3, 1, 2), acts["d1r0"]))
    a = unet.down[1]["attn"][0](ops.nchw_to_nhwc(acts["d1r0"].cuda().to(BF)), kv[0], lk)
    print("   d1a0 (oracle input) %.3e" % rel(a.permute(0, 3, 1, 2), acts["d1a0"]))
