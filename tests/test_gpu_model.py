"""GPU tests of the whole path: VQModel (product, CUDA) vs the CPU oracle restatement, same weights."""
import numpy as np
import pytest
import torch

from oracle import vit_ref, xq_oracle as xo
from test_model_cpu import small_model

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().float().cpu().numpy()


@pytest.mark.parametrize("name", ["VQ-8192", "VP2-16384", "MSVR10P2-4096", "MSBR10P2-16384"])
def test_forward_matches_oracle_fp32(name):
    """fp32, drop_path off (eval-mode ViT, quantizers in train mode so the losses are produced).
    Tolerances: latents/pixels 1e-3 relative (north star); token indices are checked bit-exact at the
    quantizer boundary, i.e. on the SAME latent the GPU produced."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    over = dict(guide_type_2="patch") if name.startswith("MSBR") else {}
    model, args = small_model(name, **over)
    model = model.cuda()
    model.train()
    model.encoder.eval(), model.decoder.eval()     # no DropPath RNG
    x = torch.rand(2, 3, 256, 256) * 2 - 1
    torch.manual_seed(5)                            # CPU generator drives dropout_rand (xqgan_model.py:274)
    dec, (vq, commit, ent, usages), sem, det, dep = model(x.cuda(), 0, 0.0, 0.0, 100)
    assert sem is None and det is None
    cfg = vit_ref.cfg_from_model_args(model.config, num_heads=6)
    ref = vit_ref.RefTokenizer(model.state_dict(), cfg)
    with torch.no_grad():
        h_gpu = model.encode(x.cuda())
        h_ref = ref.encode(x)
        np.testing.assert_allclose(npy(h_gpu), h_ref.numpy(), rtol=1e-3, atol=1e-3 * float(h_ref.abs().max()))
        # quantizer boundary on identical inputs -> indices bit-exact
        torch.manual_seed(5)
        SN = len(cfg["v_patch_nums"])
        dr = torch.randint(model.start_drop, SN + 1, (2,)).numpy() if SN > 1 else None
        quant_ref, (vq_r, cm_r, en_r) = ref.quantize(h_gpu.cpu(), dr)
        quants = model._quantizers()
        for i, hb in enumerate(ref._branches(h_gpu.cpu())):
            qm = quants[i]
            if SN == 1:
                idx_ref = xo.vq_forward(hb.numpy(), npy(qm.embedding.weight))["idx"]
                np.testing.assert_array_equal(npy(qm.last_idx), idx_ref)
            else:
                # every scale's indices, recomputed by the oracle ON THE SAME GPU LATENT of this branch
                assert len(qm.last_idx_Bl) == SN
                pn = list(cfg["v_patch_nums"])
                mods = qm.quant_resi.modules_list()
                pw = np.stack([npy(m.weight) for m in mods])
                pb = np.stack([npy(m.bias) for m in mods])
                hbn = np.ascontiguousarray(hb.numpy())
                if name.startswith("MSBR"):
                    fw = xo.lfq_forward(hbn, pw, pb, pn, using_znorm=cfg.get("codebook_l2_norm", True),
                                        codebook_drop=cfg.get("codebook_drop", 0.0), dropout=dr, scaler=npy(qm.scaler),
                                        entropy_weight=cfg.get("entropy_weight", 0.0))
                else:
                    fw = xo.vq2_forward(hbn, npy(qm.embedding.weight), pw, pb, pn, using_znorm=True,
                                        codebook_drop=cfg.get("codebook_drop", 0.0), dropout=dr)
                for si in range(SN):
                    np.testing.assert_array_equal(npy(qm.last_idx_Bl[si]).astype(np.int64), np.asarray(fw["idx"][si]))
        np.testing.assert_allclose(float(vq), float(vq_r), rtol=1e-3)
        np.testing.assert_allclose(float(commit), float(cm_r), rtol=1e-3)
        np.testing.assert_allclose(float(ent), float(en_r), rtol=1e-3, atol=1e-6)
        dec_ref = ref.decode(quant_ref)
        np.testing.assert_allclose(npy(dec), dec_ref.numpy(), rtol=1e-3, atol=1e-3 * float(dec_ref.abs().max()))
    # inference surfaces agree with each other
    model.eval()
    with torch.no_grad():
        rec = model.img_to_reconstructed_img(x.cuda())
        toks = model.img_to_idxBl(x.cuda())
        rec2 = model.decode_tokens(toks)
        assert rec.shape == (2, 3, 256, 256) and float(rec.abs().max()) <= 1.0
        np.testing.assert_allclose(npy(rec), npy(rec2), rtol=1e-5, atol=1e-5)


def test_train_step_bf16_autocast_runs_and_learns():
    model, args = small_model("VQ-8192")
    model = model.cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
    x = (torch.rand(4, 3, 256, 256) * 2 - 1).cuda()
    losses = []
    for _ in range(4):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            dec, (vq, commit, ent, usages), _, _, _ = model(x, 0, 0.0, 0.0, 100)
            loss = torch.nn.functional.mse_loss(dec.float(), x) + vq + commit
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert model.quantize.embedding.weight.grad is not None
    assert float(usages[0]) >= 0.0


def test_robusttok_perturbation_path():
    model, args = small_model("RobustTok")
    model = model.cuda().train()
    x = (torch.rand(10, 3, 256, 256) * 2 - 1).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        dec, (vq, commit, ent, usages), _, _, _ = model(x, 0, 1.0, 0.1, 100)
        (dec.float().pow(2).mean() + vq + commit).backward()
    assert torch.isfinite(model.encoder.latent_tokens.grad).all()


@pytest.mark.parametrize("name", ["VQ-8192", "MSVR10P2-4096"])
def test_pretokenize_round_trip(name, tmp_path):
    """f-2: image -> tokens -> jsonl -> tokens -> decode_tokens == img_to_reconstructed_img."""
    from imagefolder_b200 import pretokenize as pt
    model, _ = small_model(name)
    model = model.cuda().eval()
    x = torch.rand(3, 3, 256, 256) * 2 - 1
    y = torch.tensor([3, 7, 11])
    path = str(tmp_path / "tokens.jsonl")
    n = pt.pretokenize(model, [(x, y)], path, flip=True, autocast_dtype=None)
    assert n == 6
    recs = list(pt.read_tokens(path))
    assert [c for c, _ in recs] == [3, 7, 11, 3, 7, 11]
    toks = torch.stack([t for _, t in recs]).cuda()
    per_img = 256 if name == "VQ-8192" else 2 * 286
    assert toks.shape == (6, per_img)
    with torch.no_grad():
        rec = model.decode_tokens(pt.unflatten_codes(toks, model))
        ref = model.img_to_reconstructed_img(torch.cat([x, torch.flip(x, dims=[-1])]).cuda())
    np.testing.assert_allclose(npy(rec), npy(ref), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["VQ-8192", "MSVR10P2-4096"])
def test_full_training_iteration_with_loss_stack(name):
    """one iteration exactly as xqgan_train.py:448-470 wires it: tokenizer forward -> VQLoss generator branch (L2 + LPIPS +
    adaptive-weighted adversarial term through DiffAug + DINO discriminator) -> optimizer; then the discriminator branch ->
    its optimizer.  Random LPIPS / DINO weights (no network), so only the plumbing and the gradients are checked."""
    import warnings
    from imagefolder_b200.vq_loss import VQLoss
    warnings.filterwarnings("ignore", message=".*RANDOM.*")
    model, args = small_model(name)
    model = model.cuda().train()
    torch.manual_seed(1)
    vq_loss = VQLoss(disc_start=0, disc_weight=0.5, disc_type='dinodisc', disc_loss='hinge', gen_adv_loss='hinge',
                     disc_adaptive_weight=True, lecam_loss_weight=0.001, perceptual_weight=1.0, aug_prob=1.0).cuda().train()
    opt_g = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
    opt_d = torch.optim.AdamW(vq_loss.discriminator.parameters(), lr=1e-4, fused=True)
    x = (torch.rand(4, 3, 256, 256) * 2 - 1).cuda()
    head_before = [p.detach().clone() for p in vq_loss.discriminator.heads.parameters()]
    enc_before = model.encoder.latent_tokens.detach().clone()
    for step in range(2):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            recons, codebook_loss, sem_loss, detail_loss, dependency_loss = model(x, 0, 0.0, 0.0, 100)
            loss_gen = vq_loss(codebook_loss, sem_loss, detail_loss, dependency_loss, x, recons, optimizer_idx=0,
                               global_step=step + 1, last_layer=model.decoder.last_layer, fade_blur_schedule=0)
        assert torch.isfinite(loss_gen)
        opt_g.zero_grad(set_to_none=True)
        loss_gen.backward()
        assert all(p.grad is None for p in vq_loss.perceptual_loss.parameters())          # LPIPS is frozen
        assert torch.isfinite(model.decoder.last_layer.grad).all()
        opt_g.step()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss_disc = vq_loss(codebook_loss, sem_loss, detail_loss, dependency_loss, x, recons, optimizer_idx=1,
                                global_step=step + 1, fade_blur_schedule=0)
        assert torch.isfinite(loss_disc)
        opt_d.zero_grad(set_to_none=True)
        loss_disc.backward()
        opt_d.step()
    assert not torch.equal(enc_before, model.encoder.latent_tokens.detach())
    assert any(not torch.equal(a, b.detach()) for a, b in zip(head_before, vq_loss.discriminator.heads.parameters()))
    assert not any(p.requires_grad for p in vq_loss.discriminator.dino_proxy[0].parameters())   # frozen backbone


def test_cnn_encoder_decoder_match_reference_golden_on_gpu():
    """row a13 on the device: the conv Encoder / Decoder (xqgan_model.py:454-584) with the reference's weights reproduce the
    reference's outputs (fp32, TF32 off; library conv / GroupNorm / attention kernels -- no hand-written kernel on this row)."""
    from conftest import load_golden
    from imagefolder_b200.cnn import Decoder, Encoder
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = load_golden("cnn_small")
    enc = Encoder(ch=32, ch_mult=(1, 2), num_res_blocks=1, z_channels=8).eval()
    dec = Decoder(ch=32, ch_mult=(1, 2), num_res_blocks=1, z_channels=8).eval()
    enc.load_state_dict({k[4:]: torch.tensor(v) for k, v in g.items() if k.startswith("enc.")}, strict=True)
    dec.load_state_dict({k[4:]: torch.tensor(v) for k, v in g.items() if k.startswith("dec.")}, strict=True)
    enc, dec = enc.cuda(), dec.cuda()
    with torch.no_grad():
        np.testing.assert_allclose(npy(enc(torch.tensor(g["x"]).cuda())), g["h"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(npy(dec(torch.tensor(g["z"]).cuda())), g["y"], rtol=1e-3, atol=1e-4)


def test_rfid_reconstruction_loop_on_gpu():
    """f-2 remainder: the evaluation data path of xqgan_train.py:517-535 (eval mode, reconstruct, uint8 NHWC, gather) on the
    CUDA tokenizer; single process here, the world_size-2 gather is covered by tests/test_dist_cpu.py."""
    from imagefolder_b200.evaluate import reconstruct_for_fid, to_uint8_nhwc
    model, _ = small_model("MSVR10P2-4096")
    model = model.cuda().train()
    g = torch.Generator().manual_seed(3)
    batches = [(torch.rand(2, 3, 256, 256, generator=g) * 2 - 1, torch.zeros(2)) for _ in range(2)]
    smp, gt, tot = reconstruct_for_fid(model, batches)
    assert model.training and tot == 4 and smp.shape == (4, 256, 256, 3) and smp.dtype == np.uint8
    model.eval()
    with torch.no_grad():
        want = torch.cat([to_uint8_nhwc(model.img_to_reconstructed_img(x.cuda())) for x, _ in batches]).cpu().numpy()
    assert np.array_equal(smp, want)
    assert np.array_equal(gt, torch.cat([to_uint8_nhwc(x) for x, _ in batches]).numpy())
