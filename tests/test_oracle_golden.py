"""CPU: the oracle (oracle/rssm_oracle.py) against fixtures produced by the live reference."""
import numpy as np
import pytest
import torch

from oracle import rssm_oracle as O
from tests import _cases as C

torch.set_num_threads(1)


@pytest.mark.parametrize("name", C.OBSERVE_CASES)
def test_observe_matches_reference(name):
    params, x, gold, meta = C.observe_case(name)
    outs = O.observe(params, x["prev_belief"], x["prev_state"], x["actions"], x["embeds"], x["nonterms"],
                     x["eps_prior"], x["eps_post"])
    assert len(outs) == (7 if meta["use_obs"] else 4)
    keep = int(meta["keep"])
    for nm, o in zip(C.OBS_NAMES, outs):
        assert o.shape[0] == meta["T"] - 1  # T-1 steps, init sliced off (rssm.py:135)
        got = o if keep == 0 else o[-keep:]
        np.testing.assert_allclose(got.numpy(), gold[nm], rtol=2e-5, atol=2e-6, err_msg=nm)
    if meta["use_obs"]:
        kl = O.kl_sum(outs[5], outs[6], outs[2], outs[3])
        np.testing.assert_allclose(kl.numpy(), gold["kl_tb"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(O.dreamer_kl_loss(kl).item(), gold["kl_dreamer"], rtol=1e-5)
        kl_div, kl_viol, kl_loss, beta_loss = O.repo_kl_terms(kl, np.log(1e-5))
        np.testing.assert_allclose(kl_div.item(), gold["kl_mean"], rtol=1e-5)


@pytest.mark.parametrize("name", C.IMAGINE_CASES)
def test_imagine_matches_reference(name):
    params, actor, reward, value, x, gold, meta = C.imagine_case(name)
    H = int(meta["H"])
    outs = O.imagine(params, actor, x["belief"], x["state"], x["eps_action"], x["eps_prior"], H)
    for nm, o in zip(C.IMG_NAMES, outs):
        assert o.shape[0] == H - 1  # start row excluded (rssm.py:178-183)
        np.testing.assert_allclose(o.numpy(), gold[nm], rtol=2e-5, atol=2e-6, err_msg=nm)
    T1, N = outs[0].shape[:2]
    rew = O.head_forward(reward, outs[0].flatten(0, 1), outs[1].flatten(0, 1)).reshape(T1, N)
    val = O.head_forward(value, outs[0].flatten(0, 1), outs[1].flatten(0, 1)).reshape(T1, N)
    np.testing.assert_allclose(rew.numpy(), gold["rewards"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(val.numpy(), gold["values"], rtol=2e-5, atol=2e-6)
    ret = O.imagine_returns(rew, val, 0.99, 0.95)
    assert ret.shape[0] == H - 2  # dreamer.py:343-349
    np.testing.assert_allclose(ret.numpy(), gold["returns"], rtol=2e-5, atol=2e-6)


def test_lambda_return_known_answer():
    g, _ = C.load("lambda_return")
    # hand computation of the 3-step toy: gamma=0.9, lambda=0.8, bootstrap 4
    r, v, boot = [1.0, 2.0, 3.0], [0.5, 0.25, 0.125], 4.0
    nxt = v[1:] + [boot]
    inp = [r[i] + 0.9 * nxt[i] * (1 - 0.8) for i in range(3)]
    last, hand = boot, [0.0] * 3
    for i in (2, 1, 0):
        last = inp[i] + 0.9 * 0.8 * last
        hand[i] = last
    np.testing.assert_allclose(g["toy_out"].ravel(), hand, rtol=1e-6)
    out = O.lambda_return(C.t(g["toy_r"]), C.t(g["toy_v"]), 0.9 * torch.ones(3, 1), C.t(g["toy_boot"]), 0.8)
    np.testing.assert_array_equal(out.numpy(), g["toy_out"])
    out2 = O.lambda_return(C.t(g["r"]), C.t(g["v"]), 0.99 * torch.ones(13, 37), C.t(g["boot"]), 0.95)
    np.testing.assert_array_equal(out2.numpy(), g["out"])


def test_kl_closed_form():
    mp, sp, mq, sq = torch.tensor(0.3), torch.tensor(0.7), torch.tensor(-0.2), torch.tensor(1.3)
    want = np.log(1.3 / 0.7) + (0.7 ** 2 + 0.5 ** 2) / (2 * 1.3 ** 2) - 0.5
    np.testing.assert_allclose(O.kl_normal(mp, sp, mq, sq).item(), want, rtol=1e-6)
    ref = torch.distributions.kl_divergence(torch.distributions.Normal(mp, sp), torch.distributions.Normal(mq, sq))
    np.testing.assert_allclose(O.kl_normal(mp, sp, mq, sq).item(), ref.item(), rtol=1e-6)


def test_tanh_normal_entropy_matches_reference():
    g, _ = C.load("entropy_M50_A6_K100")
    ent = O.tanh_normal_entropy(C.t(g["mean"]), C.t(g["std"]), C.t(g["eps"]))
    np.testing.assert_allclose(ent.numpy(), g["entropy"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("tag", ["partial", "full", "fullwrap"])
def test_replay_indices_bit_exact(tag):
    g, _ = C.load("replay_indices")
    cap, n_push, B, L, pos, full, length = [int(v) for v in g[f"{tag}_meta"]]
    inds = O.replay_indices(g[f"{tag}_starts"], L, pos, bool(full), length)
    assert inds.shape == (L * B,)
    # rebuild the ring buffer contents the generator pushed: slot i%cap holds push index i
    owner = np.full(cap, -1, dtype=np.int64)
    for i in range(n_push):
        owner[i % cap] = i
    src = owner[inds].reshape(L, B)
    np.testing.assert_array_equal(g[f"{tag}_obs"][..., 0], src.astype(np.float32))
    np.testing.assert_array_equal(g[f"{tag}_rew"][..., 0], src.astype(np.float32))
    np.testing.assert_array_equal(g[f"{tag}_done"][..., 0], (src % 11 == 0).astype(np.float32))
    # sequences are contiguous in push order (never straddle the write head)
    assert (np.diff(src, axis=0) == 1).all()


def test_f16x3_emulation_is_within_tolerance():
    """The split-fp16 arithmetic the tcgen05 kernels use must sit well inside rtol 1e-3."""
    params, x, gold, meta = C.observe_case("observe_T8_B10_hot")
    exact = O.observe(params, x["prev_belief"], x["prev_state"], x["actions"], x["embeds"], x["nonterms"],
                      x["eps_prior"], x["eps_post"])
    emu = O.observe(params, x["prev_belief"], x["prev_state"], x["actions"], x["embeds"], x["nonterms"],
                    x["eps_prior"], x["eps_post"], prec=O.Precision("f16x3"))
    for a, b in zip(exact, emu):
        np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-4, atol=1e-5)


def test_conditional_oracle_matches_reference_fixture():
    """ConditionalTransitionModel observe / imagine (rssm.py:187-248) with a ConditionalActorModel: oracle restatement vs
    the fixture produced by the reference's own classes."""
    g, meta = C.load("conditional_N12_H7")
    seed, N, H, T, B, Cn = (int(meta[k]) for k in ("seed", "N", "H", "T", "B", "C"))
    dims = dict(O.DEFAULT_DIMS)
    pdims = dict(dims, action=dims["action"] + Cn)
    p = O.make_transition_params(seed, pdims)
    ap = O.make_mlp_params(seed + 1, dims["belief"] + dims["state"] + Cn, dims["hidden"], 2 * dims["action"], 4)
    rs = np.random.RandomState(seed + 2)
    cond = torch.from_numpy(rs.standard_normal((N, Cn)).astype(np.float32))
    x = O.make_imagine_inputs(seed + 20, N, H, dims)
    im = O.imagine_conditional(p, ap, x["belief"], x["state"], cond, x["eps_action"], x["eps_prior"], H)
    for nm, o in zip(("im_beliefs", "im_prior_states", "im_prior_means", "im_prior_std_devs"), im):
        np.testing.assert_allclose(o.numpy(), g[nm], rtol=2e-5, atol=2e-6, err_msg=nm)
    xo = O.make_observe_inputs(seed + 10, T, B, dims)
    conds = torch.from_numpy(rs.standard_normal((T - 1, B, Cn)).astype(np.float32))
    ob = O.observe(p, xo["prev_belief"], xo["prev_state"], torch.cat([xo["actions"], conds], 2), xo["embeds"], xo["nonterms"],
                   xo["eps_prior"], xo["eps_post"])
    for nm, idx in (("ob_beliefs", 0), ("ob_posterior_states", 4), ("ob_posterior_means", 5), ("ob_prior_std_devs", 3)):
        np.testing.assert_allclose(ob[idx].numpy(), g[nm], rtol=2e-5, atol=2e-6, err_msg=nm)


def test_optional_heads_match_reference_trainer_methods():
    """Oracle restatement of Dreamer.train_disag / train_inv_dynamics (dreamer.py:198-239) against the losses and
    gradients the reference's own methods produced (oracle/make_golden_heads.py)."""
    from repo_b200 import synth
    g, meta = C.load("train_heads")
    seed, T, B = int(meta["seed"]), int(meta["T"]), int(meta["B"])
    ep = {k: v.clone().requires_grad_(True) for k, v in synth.make_ensemble_params(seed, 236, 200, 200, 6).items()}
    ip = {k: v.clone().requires_grad_(True) for k, v in O.make_mlp_params(seed + 1, 430, 512, 12, 3).items()}
    x = synth.make_head_rollout(seed + 2, T, B)
    dl = O.disag_loss(ep, x["beliefs"], x["states"], x["actions"], x["nonterms"])
    il = O.inv_dyn_loss(ip, x["beliefs"], x["states"], x["actions"], x["nonterms"])
    np.testing.assert_allclose(dl.item(), g["log_disag_loss"], rtol=1e-5)
    np.testing.assert_allclose(il.item(), g["log_inv_dyn_loss"], rtol=1e-5)
    dl.backward()
    il.backward()
    for prefix, params in (("disag", ep), ("inv", ip)):
        for k, p in params.items():
            want = g[f"{prefix}_grad_{k}"]
            got = p.grad.numpy()
            np.testing.assert_allclose(np.sqrt((got.astype(np.float64) ** 2).sum()), g[f"{prefix}_gradnorm_{k}"], rtol=1e-4)
            if want.shape != got.shape:
                got = got.reshape(-1)[::97]
            np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-6 * np.abs(want).max(), err_msg=f"{prefix} {k}")


def test_disagreement_bonus_matches_reference_actor_critic():
    """dreamer.py:330-339: the ensemble-spread bonus over the imagined rollout, oracle vs the value the reference's
    train_actor_critic logged (and the other logged scalars of that call)."""
    from repo_b200 import synth
    g, meta = C.load("train_heads")
    seed, N, H = int(meta["seed"]), int(meta["N"]), int(meta["H"])
    D, S, A, Hd = 200, 30, 6, 200
    ep = synth.make_ensemble_params(seed, D + S + A, Hd, D, 6)
    tp, ap = O.make_transition_params(seed + 3), O.make_mlp_params(seed + 4, D + S, Hd, 2 * A, 4)
    rp, vp = O.make_mlp_params(seed + 5, D + S, Hd, 1, 3), O.make_mlp_params(seed + 6, D + S, Hd, 1, 3)
    y = O.make_imagine_inputs(seed + 7, N, H)
    rs = np.random.RandomState(seed + 8)
    eps_ent = torch.from_numpy(rs.standard_normal((100, (H - 1) * N, A)).astype(np.float32))
    eps_disag = torch.from_numpy(rs.standard_normal(((H - 1) * N, A)).astype(np.float32))
    b, s, pm, psd, _ = O.imagine(tp, ap, y["belief"], y["state"], y["eps_action"], y["eps_prior"], H)
    fb, fs = b.flatten(0, 1), s.flatten(0, 1)
    mean, std = O.actor_forward(ap, fb, fs)
    disag = O.disagreement(ep, fb, fs, torch.tanh(mean + std * eps_disag)).reshape(H - 1, N)
    np.testing.assert_allclose(disag.mean().item(), g["log_disagreement"], rtol=1e-4)
    ent = O.tanh_normal_entropy(mean, std, eps_ent).mean()
    np.testing.assert_allclose(ent.item(), g["log_action_entropy"], rtol=1e-4)
    rew = O.head_forward(rp, fb, fs).reshape(H - 1, N) + float(meta["disag_coef"]) * disag
    val = O.head_forward(vp, fb, fs).reshape(H - 1, N)
    ret = O.imagine_returns(rew, val)
    latent_ent = (0.5 + 0.5 * np.log(2 * np.pi) + psd.log()).sum(-1).mean()
    actor_loss = -ret.mean() - 3e-4 * ent - 0.0 * latent_ent
    np.testing.assert_allclose(actor_loss.item(), g["log_actor_loss"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(latent_ent.item(), g["log_latent_entropy"], rtol=1e-5)


def test_conv_oracle_matches_reference_fixture():
    """Oracle restatement of VisualEncoder / VisualObservationModel (encoder.py:21-41, decoder.py:28-48) against outputs of
    the reference's own modules on the same seeded weights and frames (oracle/make_golden_conv.py)."""
    from repo_b200 import synth
    g, _ = C.load("conv_stacks")
    pe, pd = synth.make_conv_params("encoder", 700), synth.make_conv_params("decoder", 701)
    x = synth.make_frames(702, 3)
    y = synth.make_imagine_inputs(703, 3, 2)
    np.testing.assert_allclose(O.visual_encoder(pe, x).numpy(), g["embed"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(O.visual_decoder(pd, y["belief"], y["state"]).numpy(), g["recon"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("algo", ["dreamer", "repo"])
def test_world_model_losses_match_reference_trainer(algo):
    """The oracle composed into the world-model loss of Dreamer.train_dynamics (dreamer.py:241-289) / RePo.train_dynamics
    (repo.py:25-96): conv encoder -> observe -> conv decoder / reward head -> reconstruction, reward and KL terms, against
    the scalars the reference's own unmodified trainers logged (oracle/make_golden_trainer.py) and their returned latents."""
    g, meta = C.load(f"train_dynamics_{algo}")
    seed, T, B = int(meta["seed"]), int(meta["T"]), int(meta["B"])
    D, S, A, Hd = 200, 30, 6, 200
    tp, rp = O.make_transition_params(seed), O.make_mlp_params(seed + 2, D + S, Hd, 1, 3)
    pe, pd = O.make_conv_params("encoder", seed + 4), O.make_conv_params("decoder", seed + 5)
    batch = O.make_train_batch(seed + 10, T, B, A)
    eps = O.make_observe_inputs(seed + 11, T, B)
    obs, actions, rewards, nonterms = batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"]
    embeds = O.visual_encoder(pe, obs.flatten(0, 1)).reshape(T, B, -1)
    outs = O.observe(tp, torch.zeros(B, D), torch.zeros(B, S), actions[:-1], embeds[1:], nonterms[:-1],
                     eps["eps_prior"], eps["eps_post"])                      # the caller's shifts: dreamer.py:256-258
    beliefs, post_s = outs[0], outs[4]
    np.testing.assert_allclose(beliefs.numpy(), g["beliefs"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(post_s.numpy(), g["posterior_states"], rtol=1e-4, atol=1e-5)
    recon = O.visual_decoder(pd, beliefs.flatten(0, 1), post_s.flatten(0, 1)).reshape(T - 1, B, 3, 64, 64)
    const = 0.5 * np.log(2 * np.pi)
    obs_loss = (0.5 * (recon - obs[1:]) ** 2 + const).sum((2, 3, 4)).mean((0, 1))            # dreamer.py:262-267
    rew = O.head_forward(rp, beliefs.flatten(0, 1), post_s.flatten(0, 1)).reshape(T - 1, B)
    reward_loss = ((0.5 * (rew - rewards[:-1].squeeze(-1)) ** 2 + const) * nonterms[:-1].squeeze(-1)).mean((0, 1))
    kl = O.kl_sum(outs[5], outs[6], outs[2], outs[3])
    np.testing.assert_allclose(obs_loss.item(), g["log_obs_loss"], rtol=1e-5)
    np.testing.assert_allclose(reward_loss.item(), g["log_reward_loss"], rtol=1e-5)
    if algo == "dreamer":
        kl_loss = O.dreamer_kl_loss(kl, float(meta["free_nats"]))
    else:
        kl_div, kl_viol, kl_loss, beta_loss = O.repo_kl_terms(kl, np.log(float(meta["init_beta"])))
        np.testing.assert_allclose(kl_div.item(), g["log_kl_div"], rtol=1e-4)
        np.testing.assert_allclose(float(beta_loss), g["log_beta_loss"], rtol=1e-4)
    np.testing.assert_allclose(float(kl_loss), g["log_kl_loss"], rtol=1e-4)
    np.testing.assert_allclose(obs_loss.item() + reward_loss.item() + float(kl_loss), g["log_model_loss"], rtol=1e-5)


def test_actor_critic_losses_match_reference_trainer():
    """The oracle composed into Dreamer.train_actor_critic (dreamer.py:304-381): imagine -> heads -> 100-sample entropy ->
    lambda-return -> actor / value losses, against the scalars the reference's unmodified method logged."""
    g, meta = C.load("train_actor_critic")
    seed, N, H = int(meta["seed"]), int(meta["N"]), int(meta["H"])
    D, S, A, Hd = 200, 30, 6, 200
    tp, ap = O.make_transition_params(seed), O.make_mlp_params(seed + 1, D + S, Hd, 2 * A, 4)
    rp, vp = O.make_mlp_params(seed + 2, D + S, Hd, 1, 3), O.make_mlp_params(seed + 3, D + S, Hd, 1, 3)
    x = O.make_imagine_inputs(seed + 20, N, H)
    eps_ent = torch.from_numpy(np.random.RandomState(seed + 30).standard_normal((100, (H - 1) * N, A)).astype(np.float32))
    b, s, pm, psd, _ = O.imagine(tp, ap, x["belief"], x["state"], x["eps_action"], x["eps_prior"], H)
    fb, fs = b.flatten(0, 1), s.flatten(0, 1)
    mean, std = O.actor_forward(ap, fb, fs)
    ent = O.tanh_normal_entropy(mean, std, eps_ent).mean()
    rew = O.head_forward(rp, fb, fs).reshape(H - 1, N)
    val = O.head_forward(vp, fb, fs).reshape(H - 1, N)
    ret = O.imagine_returns(rew, val)                                   # (H-2, N): dreamer.py:342-349
    latent_ent = (0.5 + 0.5 * np.log(2 * np.pi) + psd.log()).sum(-1).mean()
    np.testing.assert_allclose(ent.item(), g["log_action_entropy"], rtol=1e-4)
    np.testing.assert_allclose(latent_ent.item(), g["log_latent_entropy"], rtol=1e-5)
    np.testing.assert_allclose((-ret.mean() - 3e-4 * ent).item(), g["log_actor_loss"], rtol=1e-4, atol=1e-6)
    vpred = O.head_forward(vp, b[:-1].flatten(0, 1), s[:-1].flatten(0, 1)).reshape(H - 2, N)     # dreamer.py:362-368
    value_loss = (0.5 * (vpred - ret) ** 2 + 0.5 * np.log(2 * np.pi)).mean()
    np.testing.assert_allclose(value_loss.item(), g["log_value_loss"], rtol=1e-4)


def test_tia_world_model_losses_match_reference_trainer():
    """The oracle composed into TIA.train_dynamics (tia.py:87-201): task + distractor RSSMs observed on shared embeddings,
    the two 6-channel decoders mixed by the mask head, the distractor-only decoder, the adversarial distractor reward term and
    both free-nats KL terms, against the scalars the reference's unmodified TIA trainer logged."""
    import torch.nn.functional as F
    g, meta = C.load("train_dynamics_tia")
    seed, T, B = int(meta["seed"]), int(meta["T"]), int(meta["B"])
    D, S, A, Hd = 200, 30, 6, 200
    tp, dp = O.make_transition_params(seed), O.make_transition_params(seed + 6)
    rp, drp = O.make_mlp_params(seed + 2, D + S, Hd, 1, 3), O.make_mlp_params(seed + 9, D + S, Hd, 1, 3)
    pe = O.make_conv_params("encoder", seed + 4)
    pt, pdm = O.make_conv_params("decoder", seed + 5, out_channels=6), O.make_conv_params("decoder", seed + 7, out_channels=6)
    pdo, mh = O.make_conv_params("decoder", seed + 8), O.make_mask_head_params(seed + 12)
    batch = O.make_train_batch(seed + 10, T, B, A)
    eps_t, eps_d = O.make_observe_inputs(seed + 11, T, B), O.make_observe_inputs(seed + 13, T, B)
    obs, actions, rewards, nonterms = batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"]
    embeds = O.visual_encoder(pe, obs.flatten(0, 1)).reshape(T, B, -1)
    z = lambda n: torch.zeros(B, n)
    t = O.observe(tp, z(D), z(S), actions[:-1], embeds[1:], nonterms[:-1], eps_t["eps_prior"], eps_t["eps_post"])
    d = O.observe(dp, z(D), z(S), actions[:-1], embeds[1:], nonterms[:-1], eps_d["eps_prior"], eps_d["eps_post"])
    flat = lambda o: (o[0].flatten(0, 1), o[4].flatten(0, 1))
    t_recon, t_mask = O.visual_decoder(pt, *flat(t)).chunk(2, 1)                      # decoder.py:154-175
    d_recon, d_mask = O.visual_decoder(pdm, *flat(d)).chunk(2, 1)
    mask = torch.sigmoid(F.conv2d(torch.cat([t_mask, d_mask], 1), mh["0.weight"], mh["0.bias"]))   # tia.py:72, 126
    recon = (t_recon * mask + d_recon * (1 - mask)).reshape(T - 1, B, 3, 64, 64)
    const = 0.5 * np.log(2 * np.pi)
    nll = lambda x: (0.5 * (x - obs[1:]) ** 2 + const).sum((2, 3, 4)).mean((0, 1))
    obs_loss = nll(recon)
    d_obs_loss = nll(O.visual_decoder(pdo, *flat(d)).reshape(T - 1, B, 3, 64, 64))
    tgt, m = rewards[:-1].squeeze(-1), nonterms[:-1].squeeze(-1)
    rnll = lambda p_, o: ((0.5 * (O.head_forward(p_, *flat(o)).reshape(T - 1, B) - tgt) ** 2 + const) * m).mean((0, 1))
    t_reward_loss, d_reward_nll = rnll(rp, t), rnll(drp, d)
    reward_loss = t_reward_loss - 1.0 * d_reward_nll                                   # adversarial term, tia.py:152-156
    fn = float(meta["free_nats"])
    t_kl, d_kl = O.kl_sum(t[5], t[6], t[2], t[3]), O.kl_sum(d[5], d[6], d[2], d[3])
    kl_loss = torch.clamp(t_kl, min=fn).mean() + torch.clamp(d_kl, min=fn).mean()
    for name, val in (("obs_loss", obs_loss), ("d_obs_loss", d_obs_loss), ("t_reward_loss", t_reward_loss),
                      ("d_reward_loss", d_reward_nll), ("kl_loss", kl_loss), ("t_kl_div", t_kl.mean()), ("d_kl_div", d_kl.mean())):
        np.testing.assert_allclose(float(val), g["log_" + name], rtol=1e-4, err_msg=name)
    np.testing.assert_allclose(float(reward_loss), g["log_reward_loss"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(float(obs_loss + d_obs_loss + reward_loss + kl_loss), g["log_model_loss"], rtol=1e-5)
