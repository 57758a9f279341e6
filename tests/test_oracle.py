"""The oracle itself: structure the reference pins (parameter total, tuple layout, schedule), an independent
cross-check of the BERT layer restatement against HuggingFace transformers, and the committed golden vectors."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import vilbert_ref as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_parameter_count_matches_readme(full_oracle):
    """README.md:4 of the reference: "approximately ... 270 million" parameters."""
    n = R.count_parameters(full_oracle)
    assert n == 268_028_859
    assert abs(n - 270e6) / 270e6 < 0.01


def test_schedule(full_oracle):
    assert full_oracle.bert.encoder.schedule() == (
        ["T0", "T1", "T2", "T3", "T4", "T5", "C0", "T6", "V0", "C1", "T7", "V1", "C2", "T8", "V2", "C3",
         "T9", "V3", "C4", "T10", "V4", "C5", "V5", "T11"])


def test_tuple_layout_and_odd_batch(tiny_oracle):
    """Positional call and 10-tuple of worker.py:286-289; B odd -> element 3 is the [B,2] seq-relationship score."""
    cfg = tiny_oracle.config
    for B in (2, 3):
        inp = list(R.make_inputs(B, 12, 9, seed=3, vocab_size=cfg.vocab_size))
        inp[1] = inp[1][..., :cfg.v_feature_size].contiguous()
        out = tiny_oracle(*inp, output_all_attention_masks=True)
        assert len(out) == 10 and len(out[9]) == 24
        T = 13
        assert out[0].shape == (B, 200) and out[1].shape == (B, 1533) and out[2].shape == (B, 1)
        assert out[3].shape == ((B // 2, 2) if B % 2 == 0 else (B, 2))
        assert out[4].shape == (B, 3) and out[5].shape == (B, 9, cfg.v_target_size) and out[6].shape == (B, 9, 1)
        assert out[7].shape == (B, T, cfg.vocab_size) and out[8].shape == (B, T, 1)


def test_samples_are_independent(tiny_oracle):
    """No cross-sample op except the NLVR2 pairing: what makes batch sharding exact."""
    cfg = tiny_oracle.config
    inp = list(R.make_inputs(4, 12, 9, seed=4, vocab_size=cfg.vocab_size))
    inp[1] = inp[1][..., :cfg.v_feature_size].contiguous()
    whole = tiny_oracle(*inp)
    half = tiny_oracle(*[t[2:4] for t in inp])
    for i in (0, 1, 2, 4, 6, 8):
        assert torch.allclose(whole[i][2:4], half[i], atol=1e-5)
    assert torch.allclose(whole[3][1:2], half[3], atol=1e-5)       # pair (2,3) is row 1 of [B/2, 2]


def test_masks(tiny_oracle):
    """Padding tokens / masked regions must not influence valid positions; masked regions get -10000 in vision_logit."""
    cfg = tiny_oracle.config
    inp = list(R.make_inputs(2, 14, 10, seed=5, vocab_size=cfg.vocab_size, pad_regions=3))
    inp[1] = inp[1][..., :cfg.v_feature_size].contiguous()
    a = tiny_oracle(*inp)
    inp2 = [t.clone() for t in inp]
    inp2[1][:, -3:] += 100.0                       # garbage in masked regions
    pad = inp2[4] == 0
    inp2[0][pad] = 7                               # garbage ids under padding
    b = tiny_oracle(*inp2)
    assert torch.allclose(a[0], b[0], atol=1e-4)
    assert (a[6][:, -3:] < -9000).all() and (a[6][:, :-3] > -1000).all()


def test_text_layer_matches_huggingface_bert(full_oracle):
    """Independent pin of the BertLayer restatement: same weights through transformers' BertLayer."""
    tr = pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertConfig as HFConfig, BertLayer as HFLayer
    hf_cfg = HFConfig(hidden_size=768, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                      layer_norm_eps=1e-12, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    try:
        hf_cfg._attn_implementation = "eager"
    except Exception:
        pass
    hf = HFLayer(hf_cfg).eval()
    ours = full_oracle.bert.encoder.layer[3]
    missing, unexpected = hf.load_state_dict(ours.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    g = torch.Generator().manual_seed(0)
    h = torch.randn(2, 31, 768, generator=g)
    mask = torch.zeros(2, 1, 1, 31)
    mask[1, ..., 20:] = -10000.0
    with torch.no_grad():
        ref = hf(h, attention_mask=mask)
        ref = ref[0] if isinstance(ref, tuple) else ref
        out, _ = ours(h, mask)
    assert torch.allclose(out, ref, atol=2e-5), float((out - ref).abs().max())


def test_text_embeddings_match_huggingface_bert(full_oracle):
    """Independent pin of the BertEmbeddings restatement: word + position + token-type gather-sum and LayerNorm(eps 1e-12) through
    transformers' BertEmbeddings on the same weights; the task-token row ViLBERT inserts at index 1 is the only difference."""
    pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertConfig as HFConfig, BertEmbeddings as HFEmb
    hf = HFEmb(HFConfig(vocab_size=30522, hidden_size=768, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                        hidden_dropout_prob=0.0, pad_token_id=0)).eval()
    ours = full_oracle.bert.embeddings
    sd = {k: v for k, v in ours.state_dict().items() if not k.startswith("task_embeddings")}
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m or "token_type_ids" in m for m in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 30522, (3, 30), generator=g)
    seg = torch.randint(0, 2, (3, 30), generator=g)
    with torch.no_grad():
        ref = hf(input_ids=ids, token_type_ids=seg)
        out = ours(ids, seg, torch.full((3, 1), 7))
    assert out.shape == (3, 31, 768)
    assert torch.allclose(torch.cat([out[:, :1], out[:, 2:]], dim=1), ref, atol=1e-5)
    task_row = ours.LayerNorm(ours.task_embeddings(torch.full((3, 1), 7)))
    assert torch.allclose(out[:, 1:2], task_row, atol=1e-6)             # no position / type embedding on the task token


def test_co_attention_matches_torch_mha(full_oracle):
    """Independent pin of the bidirectional co-attention (arXiv 1908.02265, section 3): each direction is ordinary multi-head
    cross-attention, so torch.nn.MultiheadAttention with the module's own projection weights (kdim / vdim = the other stream's
    width) must give the same context, additive key mask included."""
    bi = full_oracle.bert.encoder.c_layer[2].biattention
    g = torch.Generator().manual_seed(2)
    v, t = torch.randn(2, 36, 1024, generator=g), torch.randn(2, 31, 768, generator=g)
    v_keep, t_keep = torch.ones(2, 36, dtype=torch.bool), torch.ones(2, 31, dtype=torch.bool)
    v_keep[1, 30:] = False
    t_keep[0, 25:] = False
    v_mask = (~v_keep).float()[:, None, None, :] * -10000.0
    t_mask = (~t_keep).float()[:, None, None, :] * -10000.0
    with torch.no_grad():
        c_text, c_img, _ = bi(v, v_mask, t, t_mask)

        def mha(q_lin, k_lin, v_lin, q_in, kv_in, keep):
            m = torch.nn.MultiheadAttention(1024, 8, bias=True, batch_first=True, kdim=kv_in.shape[-1], vdim=kv_in.shape[-1]).eval()
            # MultiheadAttention wants the query input already 1024 wide: feed it the projected query with an identity q-projection
            if m._qkv_same_embed_dim:                                 # keys / values as wide as the queries: one packed weight
                m.in_proj_weight.copy_(torch.cat([torch.eye(1024), k_lin.weight, v_lin.weight]))
            else:
                m.q_proj_weight.copy_(torch.eye(1024)); m.k_proj_weight.copy_(k_lin.weight); m.v_proj_weight.copy_(v_lin.weight)
            m.in_proj_bias.copy_(torch.cat([torch.zeros(1024), k_lin.bias, v_lin.bias]))
            m.out_proj.weight.copy_(torch.eye(1024)); m.out_proj.bias.zero_()
            out, _ = m(q_lin(q_in), kv_in, kv_in, key_padding_mask=~keep, need_weights=False)
            return out
        ref_text = mha(bi.query2, bi.key1, bi.value1, t, v, v_keep)      # text queries over image keys / values
        ref_img = mha(bi.query1, bi.key2, bi.value2, v, t, t_keep)       # image queries over text keys / values
    # -10000 additive mask vs -inf key padding: masked keys carry exp(-10000) = 0 weight in fp32 either way
    assert torch.allclose(c_text, ref_text, atol=2e-5), float((c_text - ref_text).abs().max())
    assert torch.allclose(c_img, ref_img, atol=2e-5), float((c_img - ref_img).abs().max())


def test_gelu_and_layernorm_definitions():
    x = torch.linspace(-4, 4, 101)
    assert torch.allclose(R.gelu(x), torch.nn.functional.gelu(x), atol=1e-6)          # erf form
    ln = R.BertLayerNorm(16)
    y = torch.randn(3, 16)
    assert torch.allclose(ln(y), torch.nn.functional.layer_norm(y, (16,), eps=1e-12), atol=1e-5)


def test_golden_vectors_reproduce(full_oracle):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from this oracle (seeded weights + inputs);
    they must reproduce on any host (different BLAS thread counts change only the last bits)."""
    names = {0: "vil_prediction", 1: "vil_prediction_gqa", 2: "vil_logit", 3: "vil_binary_prediction",
             4: "vil_tri_prediction", 6: "vision_logit", 8: "linguisic_logit"}
    files = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    assert len(files) >= 3
    for fn in files:
        z = np.load(os.path.join(GOLDEN, fn))
        if int(z["weight_seed"]) != 42:
            continue          # made from a real checkpoint by make_golden_from_upstream.py --checkpoint: needs that file
        # (upstream_*.npz written by make_golden_from_upstream.py on the synthetic weights are REFERENCE outputs: the day they
        # are committed this very loop pins the oracle to the reference)
        inp = R.make_inputs(int(z["B"]), int(z["Tin"]), int(z["V"]), seed=int(z["seed"]), pad_regions=int(z["pad"]))
        out = full_oracle(*inp, compute_pretraining_heads=False)
        for i, n in names.items():
            ref = torch.from_numpy(z[n])
            scale = max(1.0, float(ref.abs().max()))
            assert torch.allclose(out[i], ref, atol=2e-4 * scale), (fn, n, float((out[i] - ref).abs().max()))


def test_upstream_repin_recipe_reports_what_is_missing():
    """tests/golden/make_golden_from_upstream.py is the committed recipe that turns "parity unpinned" into reference-pinned
    fixtures once `vilbert` is importable (worker.py:44-46).  Here it is not: the script must say so and exit 3, without writing."""
    import subprocess
    import sys
    before = set(os.listdir(GOLDEN))
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, "make_golden_from_upstream.py")], capture_output=True, text=True, timeout=300)
    try:
        import vilbert.vilbert  # noqa: F401
        have = True
    except Exception:
        have = False
    if not have:
        assert r.returncode == 3, (r.returncode, r.stdout[-400:], r.stderr[-400:])
        assert "vilbert.vilbert" in r.stdout and "unpinned" in r.stdout
        assert set(os.listdir(GOLDEN)) == before


def test_synthetic_checkpoint_loads_into_oracle():
    """The product-side key list (vilbert_b200.synthetic.state_dict_spec) == the oracle's state_dict keys/shapes."""
    import vilbert_b200 as vb
    from vilbert_b200 import synthetic as S
    cfg = R.tiny_config()
    m = R.VILBertForVLTasks(cfg, num_labels=200)
    spec = S.state_dict_spec(vb.BertConfig.from_dict(cfg.to_dict()), num_labels=200)
    sd = m.state_dict()
    assert set(spec) == set(sd)
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)


def test_bf16_activation_rounding_floor(full_oracle):
    """Why the engine's default 16-bit activation format is fp16 and not bf16: rounding the GEMM operands of this
    network to bf16 (exact arithmetic otherwise) already moves the VQA logits by ~1.9e-2 > the 1e-2 the north_star
    allows, while fp16 rounding stays ~8x below it.  Pure CPU, pure oracle: this is a property of the formats."""
    inp = R.make_inputs(2, 30, 36, seed=1236)
    ref = full_oracle(*inp, compute_pretraining_heads=False)[0]
    with R.emulate_activation_rounding(full_oracle, torch.bfloat16):
        b = full_oracle(*inp, compute_pretraining_heads=False)[0]
    with R.emulate_activation_rounding(full_oracle, torch.float16):
        h = full_oracle(*inp, compute_pretraining_heads=False)[0]
    eb, eh = float((b - ref).abs().max()), float((h - ref).abs().max())
    assert 1.0e-2 < eb < 2.5e-2, eb
    assert eh < 4e-3, eh
    assert float((full_oracle(*inp, compute_pretraining_heads=False)[0] - ref).abs().max()) == 0.0   # hooks removed
