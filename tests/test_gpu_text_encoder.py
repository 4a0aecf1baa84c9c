"""-m gpu: `encode_prompt` path — B200 CLIP text encoders against the reference's own dependency,
`transformers` CLIPTextModel / CLIPTextModelWithProjection (pipeline_FaithDiff_tlc.py:605-616), evaluated in fp32 on the
same random-init weights and token ids.  Tolerance: rel-L2 <= 1.5e-2 on bf16 activations after up to 32 layers."""
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,heads,L", [(2, 12, 77), (1, 3, 16), (3, 20, 77), (2, 2, 128)])
def test_attention_causal(B, heads, L):
    from faithdiff_dev_b200 import ops
    torch.manual_seed(5)
    C = heads * 64
    qkv = torch.randn(B * L, 3 * C, device="cuda").to(BF)
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads, L, L, 0.125, causal=True)
    q, k, v = [t.float().reshape(B, L, heads, 64).transpose(1, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
    ref = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B * L, C)
    assert rel(out, ref) < 1e-2


def test_gemm_gelu_epilogue():
    from faithdiff_dev_b200 import ops
    torch.manual_seed(6)
    a = torch.randn(154, 1280, device="cuda").to(BF)
    w = (torch.randn(5120, 1280, device="cuda") / 1280 ** 0.5).to(BF)
    b = torch.randn(5120, device="cuda")
    out = ops.gemm(a, w, bias=b, act=4)
    assert rel(out, F.gelu(a.float() @ w.float().t() + b)) < 5e-3


def _models(cfg_kw, proj):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    import faithdiff_dev_b200 as fd
    torch.manual_seed(11)
    cfg = CLIPTextConfig(**cfg_kw)
    ref = (CLIPTextModelWithProjection if proj else CLIPTextModel)(cfg).eval()
    with torch.no_grad():                      # default init is tiny (std 0.02·…): make every layer matter
        for n, p_ in ref.named_parameters():
            if n.endswith("weight") and p_.dim() == 2 and "embedding" not in n:
                p_.mul_(3.0)
            if n.endswith("bias"):
                p_.normal_(0, 0.1)
    kw = dict(cfg_kw)
    if not proj:
        kw["projection_dim"] = None
    mine = fd.CLIPTextEncoder(**kw).load_state_dict(ref.state_dict())
    return ref, mine


def _ids(B, vocab, L=77, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab - 1, (B, L), generator=g)
    for b in range(B):                         # <bos> text <eos> <pad = eos>… like the CLIP tokenizer
        n = 5 + 7 * b
        ids[b, 0] = vocab - 2
        ids[b, n:] = vocab - 1
    return ids


@pytest.mark.parametrize("name,cfg_kw,proj", [
    ("small-quickgelu", dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
                             max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=96, eos_token_id=2), False),
    ("small-gelu-proj", dict(vocab_size=1000, hidden_size=256, intermediate_size=1024, num_hidden_layers=4, num_attention_heads=4,
                             max_position_embeddings=77, hidden_act="gelu", projection_dim=192, eos_token_id=2), True),
    ("small-neweos", dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                          max_position_embeddings=77, hidden_act="gelu", projection_dim=64, eos_token_id=999), True),
    ("clip-l", dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                    max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=768, eos_token_id=2), False),
])
def test_text_encoder_matches_transformers(name, cfg_kw, proj):
    ref, mine = _models(cfg_kw, proj)
    ids = _ids(2, cfg_kw["vocab_size"])
    with torch.no_grad():
        r = ref(ids, output_hidden_states=True)
    m = mine(ids, output_hidden_states=True)
    assert len(m.hidden_states) == len(r.hidden_states) == cfg_kw["num_hidden_layers"] + 1
    assert rel(m.hidden_states[0], r.hidden_states[0]) < 5e-3
    assert rel(m.hidden_states[-2], r.hidden_states[-2]) < 1.5e-2
    assert rel(m.last_hidden_state, r.last_hidden_state) < 1.5e-2
    assert rel(m[0], r[0]) < 1.5e-2                      # text_embeds (projection model) / last_hidden_state
    if not proj:
        assert rel(m.pooler_output, r.pooler_output) < 1.5e-2


class _Tok:
    """Stand-in tokenizer with the transformers call surface encode_prompt uses (pipeline…:589-595)."""
    model_max_length = 77

    def __init__(self, vocab):
        self.vocab = vocab

    def __call__(self, text, padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        text = [text] if isinstance(text, str) else text
        ids = torch.full((len(text), max_length), self.vocab - 1, dtype=torch.int64)
        for b, t in enumerate(text):
            toks = [self.vocab - 2] + [3 + (ord(ch) * 7) % (self.vocab - 10) for ch in t][:max_length - 2]
            ids[b, :len(toks)] = torch.tensor(toks)
        return types.SimpleNamespace(input_ids=ids)


def test_encode_prompt_matches_reference_logic():
    import faithdiff_dev_b200 as fd
    kw1 = dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
               max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=96, eos_token_id=2)
    kw2 = dict(vocab_size=1000, hidden_size=256, intermediate_size=1024, num_hidden_layers=4, num_attention_heads=4,
               max_position_embeddings=77, hidden_act="gelu", projection_dim=192, eos_token_id=2)
    ref1, enc1 = _models(kw1, False)
    ref2, enc2 = _models(kw2, True)
    tok = _Tok(1000)
    stub = types.SimpleNamespace(config=types.SimpleNamespace(block_out_channels=(1, 2, 3, 4), sample_size=128))
    pipe = fd.FaithDiffStableDiffusionXLPipeline(stub, None, enc1, enc2, tok, tok, stub, None, None)

    def reference(prompt, negative):
        """pipeline…:563-665 with the transformers models."""
        outs = []
        for text in (prompt, negative):
            if text is None:
                outs.append(None)
                continue
            embs = []
            for ref in (ref1, ref2):
                with torch.no_grad():
                    o = ref(tok([text]).input_ids, output_hidden_states=True)
                pooled = o[0]
                embs.append(o.hidden_states[-2])
            outs.append((torch.cat(embs, -1), pooled))
        return outs

    (pe_r, pp_r), neg = reference("a photo of a tiled cat", None)
    pe, npe, pp, npp = pipe.encode_prompt("a photo of a tiled cat")
    assert pe.shape == (1, 77, 128 + 256) and pp.shape == (1, 192) and pe.dtype == BF and pe.is_cuda
    assert rel(pe, pe_r) < 1.5e-2 and rel(pp, pp_r) < 1.5e-2
    assert torch.count_nonzero(npe) == 0 and torch.count_nonzero(npp) == 0          # force_zeros_for_empty_prompt
    (pe_r, pp_r), (ne_r, np_r) = reference("sharp, detailed", "blurry")
    pe, npe, pp, npp = pipe.encode_prompt("sharp, detailed", negative_prompt="blurry", num_images_per_prompt=2)
    assert pe.shape == (2, 77, 384) and npe.shape == (2, 77, 384) and pp.shape == (2, 192)
    assert rel(pe[1], pe_r[0]) < 1.5e-2 and rel(npe[0], ne_r[0]) < 1.5e-2 and rel(npp[1], np_r[0]) < 1.5e-2
    # tokenizer-free operation and clip_skip
    ids = tok(["sharp, detailed"]).input_ids
    pe2, _, pp2, _ = pipe.encode_prompt(input_ids=ids, num_images_per_prompt=1)
    assert torch.equal(pe2, pe[:1]) and torch.equal(pp2, pp[:1])
    pe3, _, _, _ = pipe.encode_prompt("sharp, detailed", clip_skip=1)
    with torch.no_grad():
        h1 = ref1(ids, output_hidden_states=True).hidden_states[-3]
    assert rel(pe3[..., :128], h1) < 1.5e-2
    # pre-computed embeddings pass straight through (the reference ignores them, pipeline…:1237-1243)
    a, b, c, d = pipe.encode_prompt(prompt_embeds=pe, pooled_prompt_embeds=pp, negative_prompt_embeds=npe, negative_pooled_prompt_embeds=npp)
    assert torch.equal(a, pe) and torch.equal(c, pp)
