"""B200-native CLIP text encoders for `encode_prompt`.

The reference's `encode_prompt` (FaithDiff/pipelines/pipeline_FaithDiff_tlc.py:485-717) runs two `transformers` models,
`CLIPTextModel` (CLIP ViT-L/14: 12 layers, d=768, 12 heads, quick-GELU) and `CLIPTextModelWithProjection`
(OpenCLIP bigG/14: 32 layers, d=1280, 20 heads, GELU, 1280-d text projection), over 77 token ids, and uses
`hidden_states[-2]` of both plus the projected pooled output of the second (`pipeline…:605-616`).

Both have 64-wide heads, so a layer is the kernels of the UNet's transformer blocks: LayerNorm → fused QKV GEMM (bias) →
causal flash attention (`fd_attention_d64_causal`) → out-projection GEMM (+residual) → LayerNorm → fc1 GEMM
(+quick-GELU / GELU epilogue) → fc2 GEMM (+residual).  State-dict keys are transformers' (`text_model.…`), so real
checkpoints load by name.  The call surface mirrors what the pipeline touches: `enc(input_ids, output_hidden_states=True)`
→ object with `[0]`, `.hidden_states`, `.last_hidden_state`, `.pooler_output`, `.text_embeds`; `.dtype`, `.to()`, `.config`."""
import torch

from . import ops
from .blocks import BF16, Weights

CLIP_L_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                     max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2,
                     projection_dim=None)
OPENCLIP_BIGG_CONFIG = dict(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                            num_attention_heads=20, max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5,
                            eos_token_id=2, projection_dim=1280)
_ACT = {"quick_gelu": 3, "gelu": 4}


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class TextEncoderOutput:
    """Indexable like transformers' ModelOutput: [0] is `text_embeds` for the projection model, else `last_hidden_state`."""

    def __init__(self, last_hidden_state, pooler_output, hidden_states, text_embeds):
        self.last_hidden_state, self.pooler_output = last_hidden_state, pooler_output
        self.hidden_states, self.text_embeds = hidden_states, text_embeds

    def __getitem__(self, i):
        first = self.text_embeds if self.text_embeds is not None else self.last_hidden_state
        seq = [first, self.last_hidden_state] if self.text_embeds is not None else [first, self.pooler_output]
        if self.hidden_states is not None:
            seq.append(self.hidden_states)
        return seq[i]


def eos_positions(input_ids, eos_token_id):
    """Row of the pooled token, as transformers CLIPTextTransformer.forward: configs with the legacy eos_token_id == 2
    take argmax(input_ids) (the end-of-text id 49407 is the largest id), newer ones the first occurrence of eos_token_id."""
    if eos_token_id == 2:
        return input_ids.to(torch.int64).argmax(dim=-1)
    return (input_ids.to(torch.int64) == eos_token_id).int().argmax(dim=-1)


class CLIPTextEncoder:
    def __init__(self, vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                 max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2,
                 projection_dim=None, **unused):
        if hidden_size != 64 * num_attention_heads:
            raise ValueError("CLIPTextEncoder: the B200 attention kernel needs 64-wide heads (hidden %d, heads %d)"
                             % (hidden_size, num_attention_heads))
        if hidden_act not in _ACT:
            raise ValueError("CLIPTextEncoder: unsupported hidden_act %r" % (hidden_act,))
        self.config = _Cfg(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                           num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                           max_position_embeddings=max_position_embeddings, hidden_act=hidden_act,
                           layer_norm_eps=layer_norm_eps, eos_token_id=eos_token_id, projection_dim=projection_dim)
        self.dtype, self.device = BF16, torch.device("cuda")
        self.layers = None

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def load_state_dict(self, sd, strict=True, device="cuda"):
        W = Weights(sd, device)
        c = self.config
        self.tok = W.mat("text_model.embeddings.token_embedding.weight")
        self.pos = W.mat("text_model.embeddings.position_embedding.weight")
        self.layers = []
        for i in range(c.num_hidden_layers):
            L = W.sub("text_model.encoder.layers.%d." % i)
            qkv_w = torch.cat([L.mat("self_attn.q_proj.weight"), L.mat("self_attn.k_proj.weight"), L.mat("self_attn.v_proj.weight")], 0)
            qkv_b = torch.cat([L.f32("self_attn.q_proj.bias"), L.f32("self_attn.k_proj.bias"), L.f32("self_attn.v_proj.bias")], 0)
            self.layers.append(dict(
                ln1=(L.f32("layer_norm1.weight"), L.f32("layer_norm1.bias")), qkv=(qkv_w.contiguous(), qkv_b.contiguous()),
                out=(L.mat("self_attn.out_proj.weight"), L.f32("self_attn.out_proj.bias")),
                ln2=(L.f32("layer_norm2.weight"), L.f32("layer_norm2.bias")),
                fc1=(L.mat("mlp.fc1.weight"), L.f32("mlp.fc1.bias")), fc2=(L.mat("mlp.fc2.weight"), L.f32("mlp.fc2.bias"))))
        self.final_ln = (W.f32("text_model.final_layer_norm.weight"), W.f32("text_model.final_layer_norm.bias"))
        self.proj = W.mat("text_projection.weight") if c.projection_dim else None
        unexpected = [k for k in sd if k not in W.used and not k.endswith("position_ids")]
        if strict and unexpected:
            raise KeyError("CLIPTextEncoder.load_state_dict: unexpected keys %s" % unexpected[:5])
        return self

    def __call__(self, input_ids, attention_mask=None, output_hidden_states=True, return_dict=True):
        if self.layers is None:
            raise RuntimeError("CLIPTextEncoder: load_state_dict first")
        if attention_mask is not None:
            raise NotImplementedError("CLIPTextEncoder: padding masks are not used by encode_prompt (pipeline…:605)")
        c = self.config
        ids = input_ids.to("cuda")
        B, L = ids.shape
        if L > c.max_position_embeddings or L > 128:
            raise ValueError("CLIPTextEncoder: at most %d tokens" % min(c.max_position_embeddings, 128))
        H, heads, act = c.hidden_size, c.num_attention_heads, _ACT[c.hidden_act]
        h = (self.tok[ids] + self.pos[:L][None]).reshape(B * L, H).contiguous()     # bf16 add, as the bf16 reference
        hidden = [h.view(B, L, H)]
        for ly in self.layers:
            n = ops.layernorm(h, *ly["ln1"], c.layer_norm_eps)
            qkv = ops.gemm(n, ly["qkv"][0], bias=ly["qkv"][1])
            a = ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, L, L, 0.125, causal=True)
            h = ops.gemm(a, ly["out"][0], bias=ly["out"][1], residual=h)
            n = ops.layernorm(h, *ly["ln2"], c.layer_norm_eps)
            m = ops.gemm(n, ly["fc1"][0], bias=ly["fc1"][1], act=act)
            h = ops.gemm(m, ly["fc2"][0], bias=ly["fc2"][1], residual=h)
            hidden.append(h.view(B, L, H))
        last = ops.layernorm(h, *self.final_ln, c.layer_norm_eps).view(B, L, H)
        pooled = last[torch.arange(B, device=last.device), eos_positions(ids, c.eos_token_id)]
        text_embeds = None
        if self.proj is not None:
            text_embeds = ops.gemm(pooled.contiguous(), self.proj)
        return TextEncoderOutput(last, pooled, tuple(hidden) if output_hidden_states else None, text_embeds)

    forward = __call__
