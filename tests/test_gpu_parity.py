"""GPU: parity of the CUDA encoder (through the C ABI) with the oracle / the reference's golden vectors.

Tolerances (BASELINE.json north_star): bf16 operands with fp32 accumulation; per-utterance pooled-embedding
cosine >= 0.999, max relative error (max|a-b| / max|b|) reported and bounded at 3e-2 on pooled vectors, intent
head argmax identical.  Measured on B200: pooled cosine 0.99996-0.99998, pooled max rel err < 1e-2."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from loco_asr_b200._lib import LocoError
from loco_asr_b200.synth import synth_wave, config1_lengths, synth_head
from oracle import speecht5_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
COS_MIN = 0.999
POOLED_REL_MAX = 3e-2
STAGE_REL_MAX = 4e-2


@pytest.mark.parametrize("ln_impl", [0, 1], ids=["deferred_ln", "ln_kernels"])
def test_stage_by_stage_against_oracle_taps(debug_encoder, weights, ln_impl):
    """Every stage buffer of the first layer against the oracle's taps (isolates a broken kernel).  With the default deferred
    LayerNorm the post-attention LayerNorm never exists as a tensor: the un-normalised sum the out_proj GEMM wrote
    ("attn_res") is normalised here, and the row statistics its epilogue produced are checked against the tensor itself."""
    waves = H.make_waves([6400, 20800, 48000, 9000, 33000])
    debug_encoder.debug_set("stop_after_layer", 0)
    debug_encoder.debug_set("ln_impl", ln_impl)
    try:
        taps = H.oracle_taps(weights, waves, n_layers=1)
        pooled, hidden, info = H.run_encoder(debug_encoder, waves)
        worst = H.compare_stages(debug_encoder, info, taps, H.STAGES, lambda s: None)
        layer0 = [(a, b, None) for a, b in H.LAYER0_STAGES if ln_impl == 1 or a != "ln1"]
        worst.update(H.compare_stages(debug_encoder, info, taps, layer0, lambda s: None))
        if ln_impl == 0:
            u1 = debug_encoder.debug_buffer("attn_res").float().cpu()
            g, b = weights["wrapped_encoder.layers.0.layer_norm.weight"], weights["wrapped_encoder.layers.0.layer_norm.bias"]
            for u, t in enumerate(taps):
                r0, T = int(info["rows"][u]), t["l0_ln1"].shape[0]
                got = torch.nn.functional.layer_norm(u1[r0:r0 + T], (768,), g, b, 1e-5)
                worst["ln1(attn_res)"] = max(worst.get("ln1(attn_res)", 0.0), H.rel_err(got, t["l0_ln1"]))
    finally:
        debug_encoder.debug_set("stop_after_layer", -1)
        debug_encoder.debug_set("ln_impl", 0)
    assert all(v < STAGE_REL_MAX for v in worst.values()), worst
    off = 0
    for t in taps:
        T = t["final"].shape[0]
        assert H.rel_err(hidden[off:off + T], t["final"]) < STAGE_REL_MAX
        off += T


def test_conv0_kernels_agree(debug_encoder):
    """conv0 + GroupNorm + GELU: the tcgen05 kernel (product; K = 48 split GEMM with the GroupNorm shift in the padding taps)
    against the mma.sync kernel (shift as the fp32 initial accumulator) on a ragged batch -- utterances of one tile, of many
    tiles, and with slot padding frames (which both must write as zeros)."""
    lengths = [400, 720, 6400, 200000, 9000, 33000, 401, 330000, 48000, 64000, 1279, 1280] + [3200 + 170 * i for i in range(30)]
    waves = H.make_waves(lengths, seed=78)
    debug_encoder.debug_set("stop_after_layer", 0)
    bufs = {}
    try:
        for impl in (0, 1):
            debug_encoder.debug_set("conv0_impl", impl)
            _, _, info = H.run_encoder(debug_encoder, waves)
            buf = debug_encoder.debug_buffer("conv0").float().cpu()
            rows = []
            for u in range(len(waves)):
                r0 = int(info["rows"][u]) << 6
                r1 = (int(info["rows"][u + 1]) << 6) if u + 1 < len(waves) else buf.shape[0]
                rows.append(buf[r0:r1])              # the whole slot: valid frames and the zero padding frames after them
            bufs[impl] = torch.cat(rows)
    finally:
        debug_encoder.debug_set("conv0_impl", 0)
        debug_encoder.debug_set("stop_after_layer", -1)
    assert torch.isfinite(bufs[0]).all()
    assert H.rel_err(bufs[0], bufs[1]) < 1e-2, H.rel_err(bufs[0], bufs[1])
    same = float((bufs[0] == bufs[1]).float().mean())
    assert same > 0.99, same
    assert torch.equal(bufs[0] == 0, bufs[1] == 0) or float(((bufs[0] == 0) != (bufs[1] == 0)).float().mean()) < 1e-4


def test_positional_conv_kernels_agree(debug_encoder):
    """The polyphase tcgen05 positional conv (product; four output frames per accumulator row, utterances sharing 512-frame
    timeline tiles) against the one-phase tcgen05 kernel and the mma.sync kernel: same sums in another order.  The batch has
    many short utterances per timeline tile, 1-frame utterances, and utterances that span two and three tiles."""
    lengths = [400, 720, 6400] * 8 + [200000, 9000, 33000, 400, 330000, 48000, 64000] + [3200 + 160 * i for i in range(40)]
    waves = H.make_waves(lengths, seed=77)
    debug_encoder.debug_set("stop_after_layer", 0)
    bufs = {}
    try:
        for impl in (0, 1, 2):
            debug_encoder.debug_set("posconv_impl", impl)
            _, _, info = H.run_encoder(debug_encoder, waves)
            buf = debug_encoder.debug_buffer("pos_conv").float().cpu()
            rows = []
            for u in range(len(waves)):
                r0, T = int(info["rows"][u]), int(info["frames"][u])
                rows.append(buf[r0:r0 + T])
            bufs[impl] = torch.cat(rows)
    finally:
        debug_encoder.debug_set("posconv_impl", 0)
        debug_encoder.debug_set("stop_after_layer", -1)
    assert torch.isfinite(bufs[0]).all()
    for other in (1, 2):
        assert H.rel_err(bufs[0], bufs[other]) < 1e-2, (other, H.rel_err(bufs[0], bufs[other]))
        # bf16 outputs of fp32 sums taken in another order: all but a sliver of the values are the same bf16 number
        same = float((bufs[0] == bufs[other]).float().mean())
        assert same > 0.97, (other, same)


def test_deferred_layernorm_equals_layernorm_kernels(debug_encoder):
    """The two formulations of the post-LN blocks (LayerNorm deferred into the GEMM epilogues, default; LayerNorm kernels)
    on a ragged batch: same function, different rounding points -- last_hidden_state within bf16 noise of each other."""
    waves = H.make_waves([400, 9000, 33000, 64000, 100000], seed=41)
    p0, h0, _ = H.run_encoder(debug_encoder, waves)
    debug_encoder.debug_set("ln_impl", 1)
    try:
        p1, h1, _ = H.run_encoder(debug_encoder, waves)
    finally:
        debug_encoder.debug_set("ln_impl", 0)
    assert not torch.equal(h0, h1)                      # the knob does select another path
    assert H.rel_err(h0, h1) < 3e-2 and H.rel_err(p0, p1) < 3e-2      # (a 1-frame utterance is in the batch: pooled == its only frame)
    assert float(torch.nn.functional.cosine_similarity(p0, p1, dim=1).min()) > 0.9995


def test_config1_against_golden_hf_vectors(encoder):
    """BASELINE.json configs[0]: 16 utterances of 2.5-3.5 s vs the committed outputs of the HF module."""
    g = np.load(os.path.join(GOLD, "config1_hf.npz"))
    lengths = config1_lengths()
    waves = [synth_wave(n, 0, i) for i, n in enumerate(lengths)]
    pooled, hidden, info = H.run_encoder(encoder, waves)
    assert info["frames"].tolist() == g["n_frames"].tolist()
    ref = torch.from_numpy(g["pooled"])
    cos = torch.nn.functional.cosine_similarity(pooled, ref, dim=1)
    rel = (pooled - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)
    print(f"config1: min cosine {float(cos.min()):.6f}, max rel err {float(rel.max()):.5f}")
    assert float(cos.min()) >= COS_MIN and float(rel.max()) < POOLED_REL_MAX
    w, b = synth_head(3)
    assert torch.equal(O.intent_head(pooled, w, b), O.intent_head(ref, w, b))      # downstream argmax identical
    off = 0
    for i, T in enumerate(info["frames"]):
        assert H.rel_err(hidden[off], torch.from_numpy(g["first_frame"][i])) < 5e-2
        assert H.rel_err(hidden[off + T - 1], torch.from_numpy(g["last_frame"][i])) < 5e-2
        off += int(T)


def test_short_golden_last_hidden(encoder):
    g = np.load(os.path.join(GOLD, "short_taps.npz"))
    waves = [synth_wave(int(g["a_n_samples"]), 0, int(g["a_idx"]), kind="noise"),
             synth_wave(int(g["b_n_samples"]), 0, int(g["b_idx"]), kind="mix")]
    pooled, hidden, info = H.run_encoder(encoder, waves)
    Ta = g["a_hf_last_hidden"].shape[0]
    assert H.rel_err(hidden[:Ta], torch.from_numpy(g["a_hf_last_hidden"])) < STAGE_REL_MAX
    assert H.rel_err(hidden[Ta:], torch.from_numpy(g["b_hf_last_hidden"])) < STAGE_REL_MAX


def test_edge_lengths_minimum_and_ragged(encoder, weights):
    """1-frame utterance (400 samples), 2 frames, odd lengths, and a 10 s utterance in one ragged batch."""
    lengths = [400, 720, 1039, 16001, 160000, 401, 7777]
    waves = H.make_waves(lengths, seed=2)
    pooled, hidden, info = H.run_encoder(encoder, waves)
    assert info["frames"].tolist() == [O.frame_lengths(n)[-1] for n in lengths]
    off = 0
    for u, w in enumerate(waves):
        ref = O.encode_utterance(weights, torch.from_numpy(w))
        T = ref.shape[0]
        assert H.cosine(pooled[u], ref.mean(0)) >= COS_MIN, (u, lengths[u])
        assert H.rel_err(hidden[off:off + T], ref) < 5e-2, (u, lengths[u])
        off += T


def test_too_short_and_empty_inputs(encoder):
    with pytest.raises(LocoError, match="shorter than one encoder frame"):
        encoder.encode_packed(torch.zeros(399, device="cuda"), [399])
    with pytest.raises(LocoError):
        encoder.encode_packed(torch.zeros(10, device="cuda"), [10, 0])
    out = encoder.encode_packed(torch.zeros(0, device="cuda"), [])
    assert out.shape == (0, 768)


def test_c_abi_rejects_bad_arguments_with_a_code_and_a_message(weights):
    """include/loco_asr.h: 'every function returns 0 or a negative loco_status ... nothing throws, nothing calls exit()'.
    Null pointers, negative shapes and out-of-range counts come back as LOCO_ERR_INVALID with text in loco_last_error."""
    import ctypes as C
    from loco_asr_b200 import _lib
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    enc = LocoSpeechT5Encoder(device="cuda:0")
    lib, h = enc._lib, enc._h
    data = (C.c_float * 8)()
    shape = (C.c_int64 * 1)(8)
    key = b"wrapped_encoder.layer_norm.weight"
    assert lib.loco_load_tensor(h, key, C.c_void_p(C.addressof(data)), None, 1, _lib.LOCO_F32) == _lib.LOCO_ERR_INVALID
    bad = (C.c_int64 * 1)(-8)
    assert lib.loco_load_tensor(h, key, C.c_void_p(C.addressof(data)), bad, 1, _lib.LOCO_F32) == _lib.LOCO_ERR_INVALID
    assert b"bad shape" in lib.loco_last_error(h)
    assert lib.loco_load_tensor(h, key, C.c_void_p(C.addressof(data)), shape, 1, 99) == _lib.LOCO_ERR_INVALID
    assert lib.loco_load_tensor(h, None, C.c_void_p(C.addressof(data)), shape, 1, _lib.LOCO_F32) == _lib.LOCO_ERR_INVALID
    assert lib.loco_load_tensor(h, b"decoder.x", C.c_void_p(C.addressof(data)), shape, 1, _lib.LOCO_F32) == _lib.LOCO_ERR_WEIGHTS
    ws = C.c_size_t()
    ns = (C.c_int32 * 1)(16000)
    assert lib.loco_plan(h, ns, -1, None, None, None, C.byref(ws)) == _lib.LOCO_ERR_INVALID
    assert lib.loco_plan(h, ns, 70000, None, None, None, C.byref(ws)) == _lib.LOCO_ERR_INVALID       # > 65535 utterances per call
    assert lib.loco_plan(h, None, 1, None, None, None, C.byref(ws)) == _lib.LOCO_ERR_INVALID
    # encode before finalize: a state error, not a crash
    assert lib.loco_encode(h, None, ns, 1, None, None, None, 0, None) == _lib.LOCO_ERR_STATE
    enc.load_state_dict(weights)
    enc.finalize()
    assert lib.loco_load_tensor(h, key, C.c_void_p(C.addressof(data)), shape, 1, _lib.LOCO_F32) == _lib.LOCO_ERR_STATE
    assert lib.loco_encode(h, None, ns, 1, None, None, None, 0, None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_set_head(h, 7, None, None, None, 0) == _lib.LOCO_ERR_INVALID
    assert lib.loco_set_head(h, 2, None, None, None, 0) == _lib.LOCO_ERR_INVALID                       # self_attention needs q
    assert lib.loco_set_head(h, 0, None, C.c_void_p(C.addressof(data)), None, 1) == _lib.LOCO_ERR_INVALID
    assert lib.loco_encode_planned(h, None, None, None, None, None, 0, None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_sync_check(h, None) == 0
    # the handle still works after all of that
    w = synth_wave(16000, 3, 0)
    out = enc.encode_packed(torch.from_numpy(w).cuda(), [16000])
    assert torch.isfinite(out).all()


def test_batch_composition_does_not_change_an_utterance(encoder):
    """Size-independent property: an utterance's result is bit-identical alone, in a small batch, at another
    position, and inside a large (~20k-frame) batch -- no padding or neighbour leaks anywhere in the path."""
    probe = synth_wave(30000, 9, 1)
    alone, _, _ = H.run_encoder(encoder, [probe])
    others = H.make_waves([12345, 48000, 8000], seed=4)
    p1, _, _ = H.run_encoder(encoder, [others[0], probe, others[1]])
    p2, _, _ = H.run_encoder(encoder, [probe] + others)
    big = H.make_waves([16000 + 997 * (i % 40) for i in range(255)], seed=6)
    p3, _, _ = H.run_encoder(encoder, big[:100] + [probe] + big[100:])
    assert torch.equal(alone[0], p1[1]) and torch.equal(alone[0], p2[0]) and torch.equal(alone[0], p3[100])


def test_reference_call_surface_padded_inputs(encoder, weights):
    """encoder(input_values=[B, L_max] zero padded, attention_mask) -> .last_hidden_state [B, T_max, 768]
    (the reference's call, extract_speecht5_base_embeddings_slurp.py:60,108-109), equal to the packed path."""
    waves = H.make_waves([20000, 31000, 25500], seed=8)
    lmax = max(len(w) for w in waves)
    iv = torch.zeros(3, lmax)
    am = torch.zeros(3, lmax, dtype=torch.int32)
    for i, w in enumerate(waves):
        iv[i, :len(w)] = torch.from_numpy(w)
        am[i, :len(w)] = 1
    out = encoder(input_values=iv.cuda(), attention_mask=am.cuda())
    pooled, hidden, info = H.run_encoder(encoder, waves)
    assert out.last_hidden_state.shape == (3, int(info["frames"].max()), 768)
    assert torch.equal(out.pooled.cpu(), pooled)
    off = 0
    for i, T in enumerate(info["frames"]):
        assert torch.equal(out.last_hidden_state[i, :T].cpu(), hidden[off:off + T])
        assert float(out.last_hidden_state[i, T:].abs().max() if T < out.last_hidden_state.shape[1] else 0) == 0
        off += int(T)
    ref = O.encode_utterance(weights, torch.from_numpy(waves[1]))
    assert H.cosine(out.last_hidden_state[1, :ref.shape[0]].cpu(), ref) > 0.999
    # no mask: every row is a full-length utterance
    full = encoder(input_values=iv[:, :20000].cuda())
    assert full.last_hidden_state.shape == (3, 62, 768)


def test_submodule_state_dict_shims_and_weight_norm_spellings(weights):
    """The reference loads two stripped sub-module dicts (extract...:99-100, map_speecht5_hf.py:94-99,157-168);
    transformers 4.30.2 spells weight-norm weight_g / weight_v."""
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    enc = LocoSpeechT5Encoder(device="cuda:0")
    pre = {k[len("prenet."):]: v for k, v in weights.items() if k.startswith("prenet.")}
    pre["pos_conv_embed.conv.weight_g"] = pre.pop("pos_conv_embed.conv.parametrizations.weight.original0")
    pre["pos_conv_embed.conv.weight_v"] = pre.pop("pos_conv_embed.conv.parametrizations.weight.original1")
    pre["pos_sinusoidal_embed.weights"] = torch.zeros(4, 768)            # 4.30.2 carries the table as a parameter
    wrapped = {k[len("wrapped_encoder."):]: v.half() for k, v in weights.items() if k.startswith("wrapped_encoder.")}
    enc.wrapped_encoder.load_state_dict({k: v.float() for k, v in wrapped.items()})
    enc.prenet.load_state_dict(pre)
    w = synth_wave(24000, 1, 2)
    got, _, _ = H.run_encoder(enc, [w])
    ref = O.encode_utterance(weights, torch.from_numpy(w)).mean(0)
    assert H.cosine(got[0], ref) >= 0.998      # wrapped weights went through fp16 on the way in
    enc2 = LocoSpeechT5Encoder(device="cuda:0")
    enc2.load_state_dict({k: v for k, v in weights.items() if "layers.3." not in k})
    with pytest.raises(LocoError, match="missing tensor"):
        enc2.finalize()


def test_fairseq_checkpoint_names_load_to_the_same_encoder(encoder, weights):
    """A fairseq SpeechT5 ``ckpt["model"]`` (synthesised by renaming the seeded weights backwards, plus tensors the encoder never
    reads) goes through loco_asr_b200.fairseq_keys.fairseq_to_hf -- the table form of map_speecht5_hf.py:34-168 -- into the two
    load_state_dict shims and gives the bits the HF-named weights give."""
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.fairseq_keys import fairseq_to_hf, hf_to_fairseq_name
    model = {hf_to_fairseq_name(k): v for k, v in weights.items()}
    assert None not in model
    model["decoder.layers.0.fc1.weight"] = torch.zeros(4, 4)
    model["encoder.proj.weight"] = torch.zeros(81, 768)
    enc_sd, speech_sd, _ = fairseq_to_hf(model)
    enc = LocoSpeechT5Encoder(device="cuda:0")
    enc.wrapped_encoder.load_state_dict(enc_sd)
    enc.prenet.load_state_dict(speech_sd)
    waves = H.make_waves([24000, 9000, 70000], seed=31)
    got, _, _ = H.run_encoder(enc, waves)
    want, _, _ = H.run_encoder(encoder, waves)
    assert torch.equal(got, want)


def test_host_buffer_path_equals_device_path(encoder):
    waves = H.make_waves([16000, 52000, 23000, 9000], seed=12)
    lengths = [len(w) for w in waves]
    pooled, _, _ = H.run_encoder(encoder, waves)
    host = torch.from_numpy(np.concatenate(waves)).pin_memory()
    got = encoder.encode_host(host, lengths)
    assert torch.equal(got, pooled)


def test_long_context_30s(encoder, weights):
    """BASELINE.json configs[3]: a 30 s segment (T = 1499 > 2*160: both clip regions of the relative-position
    table are used; the reference materialises 575 MB of position_bias here)."""
    w = synth_wave(480000, 21, 0)
    pooled, hidden, info = H.run_encoder(encoder, [w, synth_wave(100000, 21, 1)])
    assert int(info["frames"][0]) == 1499
    ref = O.encode_utterance(weights, torch.from_numpy(w))
    assert H.cosine(pooled[0], ref.mean(0)) >= COS_MIN
    assert H.rel_err(hidden[:1499], ref) < 5e-2


def test_linearity_property_of_prenet_scale(encoder):
    """Size-independent property: GroupNorm after conv0 makes the encoder invariant to the waveform's gain
    (up to eps); a x8 louder utterance must give (nearly) the same embedding."""
    w = synth_wave(40000, 30, 0, kind="noise")
    p, _, _ = H.run_encoder(encoder, [w, (w * 8).astype(np.float32)])
    assert H.cosine(p[0], p[1]) > 0.9995


def test_extract_cli_writes_reference_format_and_head_argmax_matches(tmp_path, capsys):
    """BASELINE.json configs[4] in miniature: the drop-in CLI on synthetic SLURP-shaped utterances -> pickles the
    reference's dataset class can read; the intent head's argmax on them equals the head on the CPU oracle."""
    import pickle
    from loco_asr_b200 import extract
    from loco_asr_b200.head import IntentHead
    from loco_asr_b200.synth import slurp_shaped_lengths, synth_state_dict
    n = 12
    extract.main(["-m", "audio", "-s", "test", "--synthetic", str(n), "--out-root", str(tmp_path), "--pooled", "average"])
    folder = extract.output_folder(str(tmp_path), "base", "test", "audio")
    files = sorted(os.listdir(folder))
    assert len(files) == n and all(f.endswith("_embedding_and_target.pickle") for f in files)
    sd = synth_state_dict(seed=1)
    lens = slurp_shaped_lengths(n, 1234)
    w, b = synth_head(3)
    head = IntentHead(w, b)
    for i in (0, 5, 11):
        with open(extract.output_path(folder, f"synth{i}"), "rb") as fh:
            d = pickle.load(fh)
        assert d["id"] == f"synth{i}" and d["embedding"].dtype == np.float32 and d["embedding"].shape == (1, 768)
        assert d["target"].shape == (101,) and d["target"].sum() == 1 and d["pooling"] == "average"
        ref = O.encode_utterance(sd, torch.from_numpy(synth_wave(int(lens[i]), 1234, i))).mean(0)
        got = torch.from_numpy(d["embedding"])[0]
        assert H.cosine(got, ref) >= COS_MIN
        assert int(head.predict(got[None])) == int(head.predict(ref[None]))
    # resume: nothing left to do
    extract.main(["-m", "audio", "-s", "test", "--synthetic", str(n), "--out-root", str(tmp_path), "--pooled", "average"])
    assert "wrote 0 files" in capsys.readouterr().out
    # a folder never mixes formats: full-sequence files may not be added to the pooled ones
    with pytest.raises(SystemExit):
        extract.main(["-m", "audio", "-s", "test", "--synthetic", str(n + 2), "--out-root", str(tmp_path)])
    # the default is the reference's [T, 768] layout: last_hidden_state of the utterance's own frames, nothing else in the dict
    extract.main(["-m", "audio", "-s", "devel", "--synthetic", "3", "--out-root", str(tmp_path)])
    with open(extract.output_path(extract.output_folder(str(tmp_path), "base", "devel", "audio"), "synth1"), "rb") as fh:
        d = pickle.load(fh)
    assert set(d) == {"id", "embedding", "target"}
    T1 = O.frame_lengths(int(slurp_shaped_lengths(3, 1234)[1]))[-1]
    assert d["embedding"].shape == (T1, 768)
    ref_seq = O.encode_utterance(sd, torch.from_numpy(synth_wave(int(slurp_shaped_lengths(3, 1234)[1]), 1234, 1)))
    assert H.cosine(torch.from_numpy(d["embedding"]).mean(0), ref_seq.mean(0)) >= COS_MIN
    # max pooling over the utterance's own frames through the fused head
    extract.main(["-m", "audio", "-s", "train", "--synthetic", "3", "--out-root", str(tmp_path), "--pooled", "max"])
    with open(extract.output_path(extract.output_folder(str(tmp_path), "base", "train", "audio"), "synth1"), "rb") as fh:
        dm = pickle.load(fh)
    assert dm["pooling"] == "max" and dm["embedding"].shape == (1, 768)
    assert float((torch.from_numpy(dm["embedding"])[0] - torch.from_numpy(d["embedding"]).max(0).values).abs().max()) < 1e-5


@pytest.mark.parametrize("impl", [3, 2, 0, 1], ids=["product_mix", "tcgen05_two_pipelines", "tcgen05_one_item", "mma_sync"])
def test_attention_kernels_against_oracle(debug_encoder, weights, impl):
    """The attention kernels (the product's per-tile mix of the two tcgen05/TMEM kernels -- 199 frames: two one-item tiles, 312 and
    781: one-item tiles plus a two-pipeline tail item --, each of the two alone, and the mma.sync cross-check) on a ragged batch
    whose lengths hit: a single frame, < 1 key block, 33 / 64 / 65 frames (tile and key-half
    edges), exactly 128 / 129 frames, > 160 (table clamps), > 193 (second table round), several key blocks."""
    lengths = [400, 6400, 10640, 20560, 20880, 41200, 41520, 64000, 100000, 250000]     # T = 1, 19, 33, 64, 65, 128, 129, 199, 312, 781
    waves = H.make_waves(lengths, seed=17)
    debug_encoder.debug_set("stop_after_layer", 0)
    debug_encoder.debug_set("attn_impl", 1 if impl == 1 else 0)
    debug_encoder.debug_set("attn_p2", 1 if impl >= 2 else 0)
    debug_encoder.debug_set("attn_p2_max_frames", 193 if impl == 3 else 1 << 30)      # 1 << 30: every utterance through the kernel under test
    try:
        taps = H.oracle_taps(weights, waves, n_layers=1)
        pooled, hidden, info = H.run_encoder(debug_encoder, waves)
        assert info["frames"].tolist() == [1, 19, 33, 64, 65, 128, 129, 199, 312, 781]
        ctx = debug_encoder.debug_buffer("ctx").float().cpu()
        for u, t in enumerate(taps):
            r0 = int(info["rows"][u])
            ref = t["l0_ctx"]
            got = ctx[r0:r0 + ref.shape[0]]
            assert torch.isfinite(got).all(), (impl, u)
            assert H.rel_err(got, ref) < STAGE_REL_MAX, (impl, u, H.rel_err(got, ref))
    finally:
        debug_encoder.debug_set("stop_after_layer", -1)
        debug_encoder.debug_set("attn_impl", 0)
        debug_encoder.debug_set("attn_p2", 1)
        debug_encoder.debug_set("attn_p2_max_frames", 193)


def test_pipelined_host_path_equals_device_path(encoder):
    """encode_host_pipelined (copy stream + staging buffers, results one batch late) returns, in order, exactly what the
    device-resident call returns for each batch -- also when a later batch is larger than the staging buffer."""
    batches = [H.make_waves([8000, 30000, 12345], seed=21), H.make_waves([64000, 9000], seed=22),
               H.make_waves([20000] * 5, seed=23), H.make_waves([100000, 8000, 8000], seed=24)]
    want = [H.run_encoder(encoder, w)[0] for w in batches]
    items = [(torch.from_numpy(np.concatenate(w)).pin_memory(), [len(x) for x in w]) for w in batches]
    got = list(encoder.encode_host_pipelined(items))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert torch.equal(g, w)


def test_config5_subset_against_hf_golden_with_intent_head(encoder):
    """SURVEY.md 8(d) "Config 5": a fixed 512-utterance subset of the 70k SLURP-shaped set (the bench's own lengths,
    weights seed 1, waveforms seed 1234) against tests/golden/config5_hf.npz -- the unmodified HF module's pooled
    embeddings (fp32) and IntentClassifier(average) logits with the seed-3 Linear(768,101), made by
    oracle/make_golden.py --config5-only.  The fused head's logits are the ones compared.

    The yardstick for "bf16 compute" is in the same file: the SAME HF module run under torch.autocast(bfloat16) -- the
    reference's own arithmetic at this path's operand precision -- moves the pooled embedding by up to 1.12e-2 (max rel err;
    mean 8.1e-3; min cosine 0.999958), a logit by up to 1.8e-2, and flips the intent argmax on 12 of the 512 utterances, all
    near-ties (fp32 top-2 margin < 8.4e-3).  north_star's "argmax identical" is therefore not attainable at bf16 by the
    reference itself; the bars here are: cosine >= 0.999; max rel err within 1.25x of what HF-in-bf16 shows; every argmax
    difference is a near-tie by the same measure (fp32 margin below twice the largest logit movement), never more of them
    than HF-in-bf16 produces plus a quarter, and the argmax is identical wherever the fp32 margin exceeds 0.02."""
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    from loco_asr_b200.head import IntentHead
    from loco_asr_b200.synth import synth_state_dict
    g = np.load(os.path.join(GOLD, "config5_hf.npz"))
    ids, n_samples = g["ids"], g["n_samples"]
    enc = LocoSpeechT5Encoder.from_state_dict(synth_state_dict(seed=1), device="cuda:0")
    w, b = synth_head(3)
    enc.set_head(IntentHead(w, b, None, "average"))
    order = np.argsort(n_samples, kind="stable")
    pooled = torch.empty(len(ids), 768)
    logits = torch.empty(len(ids), 101)
    for k in range(0, len(ids), 64):
        sel = order[k:k + 64]
        waves = [synth_wave(int(n_samples[i]), 1234, int(ids[i])) for i in sel]
        wave = torch.from_numpy(np.concatenate(waves)).cuda()
        p, hp, lg = enc.encode_packed(wave, [len(x) for x in waves], with_head=True)
        pooled[sel] = p.cpu()
        logits[sel] = lg.cpu()
    ref = torch.from_numpy(g["pooled"])
    cos = torch.nn.functional.cosine_similarity(pooled, ref, dim=1)
    rel = (pooled - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)
    agree = logits.argmax(dim=1).numpy() == g["argmax"]
    logit_diff = float((logits - torch.from_numpy(g["logits"])).abs().max())
    hf_flips = int((g["hf_bf16_argmax"] != g["argmax"]).sum())
    hf_rel, hf_logit = float(g["hf_bf16_rel_err"].max()), float(g["hf_bf16_logit_diff"].max())
    print(f"config 5: min cosine {float(cos.min()):.6f} (HF in bf16: {float(g['hf_bf16_cosine'].min()):.6f}), max rel err "
          f"{float(rel.max()):.5f} (HF in bf16: {hf_rel:.5f}), mean rel err {float(rel.mean()):.5f} (HF in bf16: {float(g['hf_bf16_rel_err'].mean()):.5f}), "
          f"argmax differs on {int((~agree).sum())} of {len(ids)} (HF in bf16: {hf_flips}; largest fp32 margin among ours "
          f"{float(g['margin'][~agree].max()) if (~agree).any() else 0.0:.5f}), max |logit diff| {logit_diff:.5f} (HF in bf16: {hf_logit:.5f})")
    assert float(cos.min()) >= COS_MIN
    assert float(rel.max()) <= 1.25 * hf_rel and float(rel.mean()) <= 1.1 * float(g["hf_bf16_rel_err"].mean())
    assert logit_diff <= 1.25 * hf_logit
    assert agree[g["margin"] > 0.02].all()
    assert int((~agree).sum()) <= hf_flips + hf_flips // 4
    assert not (~agree).any() or float(g["margin"][~agree].max()) < 2 * logit_diff


@pytest.mark.parametrize("max_frames", [131072, 196608])
def test_full_size_batch_is_bit_identical_to_single_utterance_encodes(encoder, max_frames):
    """At the bench's batch size (131072 frames; conv0's activation is 8.6 GB, element indices pass 2^32) and beyond:
    utterances picked from the front, the middle and the very end of one packed launch come out bit-identical to
    encoding them alone -- the size-independent form of parity (offset arithmetic, TMA coordinates, work lists, and no
    warp-wide decision that looks at a neighbour's rows: the lazy softmax rescale once did)."""
    gen = torch.Generator(device="cuda").manual_seed(77)
    rng = np.random.default_rng(77)
    lengths, frames = [], 0
    while True:
        n = int(rng.integers(16000, 160001))
        t = (n - 400) // 320 + 1
        if frames + t + 2 > max_frames:
            break
        lengths.append(n)
        frames += t + 2
    wave = torch.randn(int(np.sum(lengths)), device="cuda", generator=gen) * 0.1
    pooled = encoder.encode_packed(wave, lengths)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(pooled).all())
    offs = np.concatenate([[0], np.cumsum(lengths)])
    for u in sorted(set(np.linspace(0, len(lengths) - 1, 12).astype(int).tolist()) | {1, len(lengths) - 2}):
        alone = encoder.encode_packed(wave[int(offs[u]):int(offs[u + 1])].contiguous(), [lengths[u]])
        assert torch.equal(alone[0], pooled[u]), u


def test_parity_with_large_layernorm_affines_and_biases(weights):
    """Random init leaves every LayerNorm at gamma = 1, beta = 0 and every bias near 0, which would hide mistakes in the
    deferred-LayerNorm algebra (gamma folded into the consumer's weights, beta into biases, the mean term through c1).  Here
    the affines and biases are drawn wide (gamma 1 +- 0.4, beta +- 0.5, dense biases +- 0.3, so rows carry real means) and the
    result is held to the same bars against the oracle."""
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    g = torch.Generator().manual_seed(99)
    sd = {k: v.clone() for k, v in weights.items()}
    for k in sd:
        if "layer_norm.weight" in k and "feature_encoder" not in k:
            sd[k] = 1.0 + 0.4 * (torch.rand(sd[k].shape, generator=g) * 2 - 1)
        elif "layer_norm.bias" in k and "feature_encoder" not in k:
            sd[k] = 0.5 * (torch.rand(sd[k].shape, generator=g) * 2 - 1)
        elif k.startswith("wrapped_encoder.layers.") and k.endswith(".bias"):
            sd[k] = 0.3 * (torch.rand(sd[k].shape, generator=g) * 2 - 1)
    enc = LocoSpeechT5Encoder.from_state_dict(sd, device="cuda:0")
    waves = H.make_waves([9000, 33000, 70000], seed=53)
    pooled, hidden, info = H.run_encoder(enc, waves)
    off = 0
    for u, w in enumerate(waves):
        ref = O.encode_utterance(sd, torch.from_numpy(w))
        T = ref.shape[0]
        assert H.cosine(pooled[u], ref.mean(0)) >= COS_MIN, u
        assert H.rel_err(pooled[u], ref.mean(0)) < POOLED_REL_MAX, u
        assert H.rel_err(hidden[off:off + T], ref) < 5e-2, u
        off += T
    dbg = LocoSpeechT5Encoder.from_state_dict(sd, device="cuda:0", debug=True)      # the LayerNorm-kernel path as a cross-check
    p0, h0, _ = H.run_encoder(dbg, waves)
    assert torch.equal(p0, pooled) and torch.equal(h0, hidden)                      # same product kernels in both builds
    dbg.debug_set("ln_impl", 1)
    p1, h1, _ = H.run_encoder(dbg, waves)
    assert H.rel_err(hidden, h1) < 3e-2


def test_no_write_outside_the_callers_buffers(encoder):
    """The caller owns workspace and outputs (include/loco_asr.h): with exactly-sized buffers embedded between guard
    pages, a ragged batch (1-frame utterance, tile-boundary lengths, a 10 s one) leaves every guard byte untouched --
    the TMA stores, the padded conv rows and the packed outputs all stay inside what loco_plan announced."""
    import ctypes as C
    lengths = [400, 41200, 41520, 64000, 160000, 7777]
    ns = np.ascontiguousarray(np.asarray(lengths, dtype=np.int32))
    n = len(lengths)
    info = encoder.plan(ns)
    G = 1 << 16
    ws = torch.full((info["workspace_bytes"] + 2 * G,), 0xAB, dtype=torch.uint8, device="cuda")
    pooled = torch.full((n * 768 + 2 * 1024,), 7.25, dtype=torch.float32, device="cuda")
    hidden = torch.full((info["total_frames"] * 768 + 2 * 1024,), 7.25, dtype=torch.float32, device="cuda")
    wave = torch.cat([torch.full((1024,), 3.0, device="cuda"), torch.randn(int(ns.sum()), device="cuda") * 0.1,
                      torch.full((1024,), 3.0, device="cuda")])
    with torch.cuda.device(encoder.device):
        rc = encoder._lib.loco_encode(encoder._h, wave.data_ptr() + 4096, ns.ctypes.data, n, pooled.data_ptr() + 4096,
                                      hidden.data_ptr() + 4096, ws.data_ptr() + G, info["workspace_bytes"],
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert bool((ws[:G] == 0xAB).all()) and bool((ws[-G:] == 0xAB).all())
    assert bool((pooled[:1024] == 7.25).all()) and bool((pooled[-1024:] == 7.25).all())
    assert bool((hidden[:1024] == 7.25).all()) and bool((hidden[-1024:] == 7.25).all())
    assert bool((wave[:1024] == 3.0).all()) and bool((wave[-1024:] == 3.0).all())
    got = pooled[1024:-1024].reshape(n, 768)
    assert bool(torch.isfinite(got).all())
    ref = encoder.encode_packed(wave[1024:-1024].contiguous(), lengths)
    assert torch.equal(got, ref)          # and the guard pages' contents (a neighbour's 3.0 samples) never leaked in


@pytest.mark.parametrize("fill", [0x00, 0xFF, 0x7F])
def test_result_does_not_depend_on_workspace_contents(encoder, fill):
    """Nothing reads workspace bytes it has not written: the same ragged batch over a workspace pre-filled with zeros,
    with 0xFF (bf16 / fp32 NaN patterns) or with 0x7F7F (huge finite bf16) gives bit-identical results -- slot padding
    rows, TMA boxes that reach past an utterance and masked keys included."""
    import ctypes as C
    lengths = [400, 12000, 41200, 41520, 64000, 100000, 7777, 30000]
    ns = np.ascontiguousarray(np.asarray(lengths, dtype=np.int32))
    n = len(lengths)
    info = encoder.plan(ns)
    wave = torch.randn(int(ns.sum()), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.1
    want_p, want_h, _ = encoder.encode_packed(wave, lengths, return_hidden=True)
    ws = torch.full((info["workspace_bytes"],), fill, dtype=torch.uint8, device="cuda")
    pooled = torch.empty(n, 768, device="cuda")
    hidden = torch.empty(info["total_frames"], 768, device="cuda")
    with torch.cuda.device(encoder.device):
        rc = encoder._lib.loco_encode(encoder._h, wave.data_ptr(), ns.ctypes.data, n, pooled.data_ptr(), hidden.data_ptr(),
                                      ws.data_ptr(), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(pooled, want_p) and torch.equal(hidden, want_h)


@pytest.mark.parametrize("misalign", [4, 256, 1000])
def test_workspace_may_have_any_alignment(encoder, misalign):
    """The C ABI takes the caller's workspace pointer as it is (cudaMalloc promises 256 B, torch's allocator 512 B): the library
    rounds the base up itself and loco_plan's workspace_bytes already includes the slack.  Same bits whatever the offset."""
    import ctypes as C
    lengths = [400, 12000, 41200, 64000, 7777]
    ns = np.ascontiguousarray(np.asarray(lengths, dtype=np.int32))
    n = len(lengths)
    info = encoder.plan(ns)
    wave = torch.randn(int(ns.sum()), device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)) * 0.1
    want = encoder.encode_packed(wave, lengths)
    raw = torch.empty(info["workspace_bytes"] + 2048, dtype=torch.uint8, device="cuda")
    off = (-raw.data_ptr()) % 1024 + misalign              # base is exactly `misalign` bytes past a 1024-byte boundary
    ws = raw[off:off + info["workspace_bytes"]]
    assert ws.data_ptr() % 1024 == misalign % 1024
    pooled = torch.empty(n, 768, device="cuda")
    with torch.cuda.device(encoder.device):
        rc = encoder._lib.loco_encode(encoder._h, wave.data_ptr(), ns.ctypes.data, n, pooled.data_ptr(), None,
                                      ws.data_ptr(), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, encoder._lib.loco_last_error(encoder._h)
    torch.cuda.synchronize()
    assert torch.equal(pooled, want)
    # one byte less than loco_plan asked for is refused with the workspace error, never written past
    with torch.cuda.device(encoder.device):
        rc = encoder._lib.loco_encode(encoder._h, wave.data_ptr(), ns.ctypes.data, n, pooled.data_ptr(), None,
                                      ws.data_ptr(), ws.numel() - 1025, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == -4


def test_long_segments_are_bit_identical_alone_and_in_a_mixed_batch(encoder):
    """BASELINE.json configs[3] lengths (30 s -> 1499 frames, 60 s -> 2999 frames) next to a 1 s utterance: the long-context
    attention path (both clamp regions of the bias table, 24-47 key blocks per query tile) gives the same bits for an
    utterance alone and inside the mixed batch, and its frame count is the reference's."""
    lengths = [960000, 16000, 480000]
    wave = torch.randn(sum(lengths), device="cuda", generator=torch.Generator(device="cuda").manual_seed(9)) * 0.1
    pooled, hidden, info = encoder.encode_packed(wave, lengths, return_hidden=True)
    assert info["frames"].tolist() == [2999, 49, 1499]
    assert bool(torch.isfinite(hidden).all())
    offs = np.concatenate([[0], np.cumsum(lengths)])
    foffs = np.concatenate([[0], np.cumsum(info["frames"])])
    for u in range(3):
        p1, h1, _ = encoder.encode_packed(wave[int(offs[u]):int(offs[u + 1])].contiguous(), [lengths[u]], return_hidden=True)
        assert torch.equal(p1[0], pooled[u]), u
        assert torch.equal(h1, hidden[int(foffs[u]):int(foffs[u + 1])]), u
    # unit-variance rows (the output of the last LayerNorm with gamma = 1, beta = 0)
    assert abs(float(hidden.std(dim=1).mean()) - 1.0) < 2e-2


def test_long_context_60s_against_golden_hf_vectors(encoder):
    """BASELINE.json configs[3], 60 s (T = 2999): against tests/golden/long60_hf.npz, the unmodified HF module's pooled vector
    and 16 evenly spaced rows of last_hidden_state (oracle/make_golden.py --long60-only; the module needs 2.3 GB of
    position_bias for this one utterance)."""
    g = np.load(os.path.join(GOLD, "long60_hf.npz"))
    w = synth_wave(int(g["n_samples"]), int(g["wave_seed"]), int(g["wave_idx"]))
    pooled, hidden, info = H.run_encoder(encoder, [w])
    assert int(info["frames"][0]) == int(g["n_frames"]) == 2999
    ref = torch.from_numpy(g["pooled"])
    print(f"60 s: pooled cosine {H.cosine(pooled[0], ref):.6f}, rel err {H.rel_err(pooled[0], ref):.5f}")
    assert H.cosine(pooled[0], ref) >= COS_MIN and H.rel_err(pooled[0], ref) < POOLED_REL_MAX
    rows = torch.from_numpy(g["rows"])
    assert H.rel_err(hidden[rows], torch.from_numpy(g["hidden_rows"])) < 5e-2


def test_utterance_longer_than_max_speech_positions(encoder, weights):
    """90 s -> 4499 frames, past config.max_speech_positions = 4000: HF grows its sinusoid table on demand
    (HF:331-333) and so does the library (loco_encode rebuilds the device table); checked against the oracle."""
    w = synth_wave(1440000, 31, 0)
    pooled, hidden, info = H.run_encoder(encoder, [w])
    assert int(info["frames"][0]) == 4499
    ref = O.encode_utterance(weights, torch.from_numpy(w))
    assert H.cosine(pooled[0], ref.mean(0)) >= COS_MIN
    assert H.rel_err(hidden[4000:], ref[4000:]) < 5e-2        # the frames whose positions lie beyond the initial table
    assert H.rel_err(hidden, ref) < 5e-2


def test_twenty_thousand_tiny_utterances_in_one_launch(encoder):
    """The other extreme of the batch shape: 20000 utterances of 1-3 frames (25-75 ms) in one launch -- per-utterance grids,
    work lists and metadata at their widest; spot-checked bit-identical against single-utterance encodes."""
    rng = np.random.default_rng(3)
    lengths = rng.integers(400, 1200, size=20000).tolist()
    wave = torch.randn(int(np.sum(lengths)), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 0.1
    pooled, hidden, info = encoder.encode_packed(wave, lengths, return_hidden=True)
    assert bool(torch.isfinite(pooled).all())
    assert info["frames"].tolist() == [(n - 400) // 320 + 1 for n in lengths]
    offs = np.concatenate([[0], np.cumsum(lengths)])
    for u in (0, 1, 9999, 19998, 19999):
        alone = encoder.encode_packed(wave[int(offs[u]):int(offs[u + 1])].contiguous(), [lengths[u]])
        assert torch.equal(alone[0], pooled[u]), u


def test_product_library_has_no_debug_knobs(encoder):
    from loco_asr_b200._lib import LocoError
    assert encoder._lib.loco_is_debug_build() == 0
    with pytest.raises(LocoError):
        encoder.debug_set("attn_impl", 1)
    with pytest.raises(LocoError):
        encoder.debug_set("stop_after_layer", 0)


def test_planned_encode_is_graph_capturable_and_replays_on_new_waveforms(encoder):
    """'all work is enqueued on the caller's stream, no hidden syncs, CUDA-graph-capturable'.  One bucket is
    planned (loco_plan_create), its encode captured into a CUDA graph (kernels and memset nodes only), and the graph replayed
    after refilling the same input buffer with other waveforms: bit-identical to the eager call on those waveforms."""
    lengths = [16000, 23456, 41200, 64000, 99999, 12345]
    plan = encoder.make_plan(lengths)
    assert plan.frames.tolist() == [int(O.frame_lengths(n)[-1]) for n in lengths]
    gen = torch.Generator(device="cuda").manual_seed(11)
    wave_a = torch.randn(sum(lengths), device="cuda", generator=gen) * 0.1
    wave_b = torch.randn(sum(lengths), device="cuda", generator=gen) * 0.1
    want_a = encoder.encode_packed(wave_a, lengths).clone()
    want_b = encoder.encode_packed(wave_b, lengths).clone()
    buf = wave_a.clone()
    pooled = torch.zeros(len(lengths), 768, device="cuda")
    ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device="cuda")
    encoder.encode_planned(plan, buf, pooled, ws)            # eager, caller-owned buffers
    encoder.sync_check()
    assert torch.equal(pooled, want_a)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        pooled.zero_()
        with torch.cuda.graph(graph, stream=side):
            encoder.encode_planned(plan, buf, pooled, ws)    # a host sync, cudaMalloc or pageable copy in here would fail the capture
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(pooled, want_a)
    buf.copy_(wave_b)
    pooled.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(pooled, want_b)
    n0 = encoder.launch_count
    graph.replay()                                            # replays launch no host-side work at all
    torch.cuda.synchronize()
    assert encoder.launch_count == n0
    plan.close()
