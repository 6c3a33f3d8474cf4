"""CPU-side checks of the Gemma-2 support: config parsing, weight naming and
fusion, the ctypes view of the extended model config."""
import torch


def test_gemma2_builtin_spec_roundtrip():
    from llmq_b200.model import BUILTIN_SPECS, ModelSpec
    s = BUILTIN_SPECS["gemma-2-9b"]
    r = ModelSpec.from_hf_config(s.to_hf_config(), name="x")
    assert (r.arch, r.head_dim, r.sliding_window, r.attn_softcap, r.final_softcap, r.tie_embeddings) == \
        ("gemma2", 256, 4096, 50.0, 30.0, True)
    assert abs(r.attn_scale - 1 / 16) < 1e-9 and r.embed_scale == 59.75  # bf16(sqrt(3584))
    # SURVEY §8(d)-style bookkeeping: bytes streamed per decode step, KV bytes per token
    assert s.weight_bytes_per_step() == 2 * (42 * (8192 * 3584 + 3584 * 4096 + 3 * 14336 * 3584) + 256000 * 3584)
    assert s.kv_bytes_per_token() == 2 * 42 * 8 * 256 * 2
    assert BUILTIN_SPECS["llama-3-8b"].arch == "llama" and BUILTIN_SPECS["llama-3-8b"].embed_scale == 0.0


def test_gemma2_hf_config_variants():
    from llmq_b200.model import ModelSpec
    import pytest
    cfg = {"architectures": ["Gemma2ForCausalLM"], "hidden_size": 2304, "num_hidden_layers": 26,
           "num_attention_heads": 8, "num_key_value_heads": 4, "head_dim": 256, "intermediate_size": 9216,
           "vocab_size": 256000, "rms_norm_eps": 1e-6, "query_pre_attn_scalar": 256,
           "attn_logit_softcapping": 50.0, "final_logit_softcapping": 30.0, "sliding_window": 4096,
           "hidden_activation": "gelu_pytorch_tanh", "eos_token_id": [1, 107], "bos_token_id": 2,
           "layer_types": ["sliding_attention", "full_attention"] * 13}
    s = ModelSpec.from_hf_config(cfg, "gemma-2-2b-it")
    assert s.arch == "gemma2" and s.eos_token_id == 1 and s.rope_theta == 10000.0 and s.tie_embeddings
    with pytest.raises(ValueError):
        ModelSpec.from_hf_config({**cfg, "layer_types": ["full_attention", "sliding_attention"] * 13}, "x")
    with pytest.raises(ValueError):
        ModelSpec.from_hf_config({**cfg, "hidden_activation": "gelu"}, "x")
    with pytest.raises(ValueError):
        ModelSpec.from_hf_config({"architectures": ["Qwen3ForCausalLM"]}, "x")


def test_gemma2_weight_fusion_names():
    from llmq_b200.model import ModelSpec, fuse_hf_weights
    from oracle.gemma2 import Gemma2Dims, random_gemma2_weights
    d = Gemma2Dims(hidden=256, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64, intermediate=512, vocab=512)
    spec = ModelSpec(hidden=256, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64, intermediate=512, vocab=512,
                     tie_embeddings=True, arch="gemma2")
    fused = dict(fuse_hf_weights(spec, random_gemma2_weights(d, seed=1)))
    assert "lm_head" not in fused and "layers.0.post_norm" not in fused
    for i in range(2):
        for n in ("input_norm", "post_attn_norm", "pre_ffn_norm", "post_ffn_norm"):
            assert fused[f"layers.{i}.{n}"].shape == (1, 256)
        assert fused[f"layers.{i}.qkv"].shape == (8 * 64, 256) and fused[f"layers.{i}.gate_up"].shape == (1024, 256)


def test_model_config_struct_carries_arch_fields():
    from llmq_b200 import lib
    c = lib.ModelConfig(hidden=256, arch=lib.ARCH_GEMMA2, sliding_window=4096, attn_softcap=50.0,
                        final_softcap=30.0, embed_scale=16.0)
    assert (c.arch, c.sliding_window, c.attn_softcap, c.final_softcap, c.embed_scale) == (1, 4096, 50.0, 30.0, 16.0)
    h = lib.load()
    bad = lib.ModelConfig(hidden=256, n_layers=1, n_q_heads=4, n_kv_heads=1, head_dim=64, intermediate=256,
                          vocab=256, block_size=16, max_tokens=16, max_seqs=4, max_pos=64, rms_eps=1e-5,
                          attn_scale=0.125, arch=lib.ARCH_LLAMA, attn_softcap=50.0)
    assert h.b200q_model_workspace_bytes(bad) == -1  # llama takes no soft-capping: rejected by check_cfg
    ok = lib.ModelConfig(hidden=256, n_layers=1, n_q_heads=4, n_kv_heads=1, head_dim=256, intermediate=256,
                         vocab=256, block_size=16, max_tokens=16, max_seqs=4, max_pos=64, rms_eps=1e-6,
                         attn_scale=0.0625, arch=lib.ARCH_GEMMA2, attn_softcap=50.0, final_softcap=30.0,
                         sliding_window=32, embed_scale=16.0)
    assert h.b200q_model_workspace_bytes(ok) > 0


def test_gemma2_model_dir_is_a_real_hf_checkpoint(tmp_path):
    """write_model_dir(gemma2 spec) -> transformers loads it as Gemma2ForCausalLM; its logits equal
    the oracle's on the same seeded weights; resolve_model / load_hf_state_dict / fuse_hf_weights
    (the product's loading path up to the device) see the same tensors."""
    from transformers import AutoConfig, AutoModelForCausalLM

    from llmq_b200.fixtures import seeded_state_dict, write_model_dir
    from llmq_b200.model import ModelSpec, fuse_hf_weights, interleave_gate_up, load_hf_state_dict, resolve_model
    from oracle.gemma2 import Gemma2Dims, Gemma2Oracle

    spec = ModelSpec(hidden=256, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64, intermediate=512, vocab=1024,
                     rms_eps=1e-6, rope_theta=10000.0, tie_embeddings=True, max_position_embeddings=256,
                     name="tiny-gemma2", arch="gemma2", query_pre_attn_scalar=64.0, attn_softcap=50.0,
                     final_softcap=30.0, sliding_window=16)
    mdir = write_model_dir(str(tmp_path / "tiny-gemma2"), spec, seed=11)
    cfg = AutoConfig.from_pretrained(mdir)
    assert cfg.model_type == "gemma2" and cfg.sliding_window == 16 and cfg.attn_logit_softcapping == 50.0
    got_spec, got_dir = resolve_model(mdir)
    assert got_dir == mdir
    for f in ("arch", "hidden", "n_layers", "n_q_heads", "n_kv_heads", "head_dim", "intermediate", "vocab",
              "sliding_window", "attn_softcap", "final_softcap", "query_pre_attn_scalar", "tie_embeddings"):
        assert getattr(got_spec, f) == getattr(spec, f), f
    sd = seeded_state_dict(spec, 11)
    fused = dict(fuse_hf_weights(got_spec, load_hf_state_dict(mdir)))
    assert torch.equal(fused["layers.1.pre_ffn_norm"][0], sd["model.layers.1.pre_feedforward_layernorm.weight"])
    assert torch.equal(fused["layers.0.gate_up"], interleave_gate_up(sd["model.layers.0.mlp.gate_proj.weight"],
                                                                      sd["model.layers.0.mlp.up_proj.weight"]))
    hf = AutoModelForCausalLM.from_pretrained(mdir, torch_dtype=torch.float32, attn_implementation="eager").eval()
    ids = torch.randint(3, 1000, (1, 40), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = hf(ids).logits[0]
    d = Gemma2Dims(hidden=256, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64, intermediate=512, vocab=1024,
                   query_pre_attn_scalar=64.0, sliding_window=16, max_pos=256)
    mine, _ = Gemma2Oracle(d, sd, "fp32").forward(ids[0], torch.arange(40))
    assert (mine - ref).abs().max().item() < 5e-5
