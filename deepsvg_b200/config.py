"""`model_cfg` classes: the attribute names and defaults are the reference's public contract
(deepsvg/model/config.py:4-108).  `SVGTransformer` / `SVGLoss` accept ANY object carrying these attributes, so
the reference's own config instances (e.g. configs/deepsvg/hierarchical_ordered.py:4-9) plug in unchanged; these
mirrors exist because /root/reference does not travel to the GPU box.
"""

N_COMMANDS = 7  # len(SVGTensor.COMMANDS_SIMPLIFIED): m, l, c, a, EOS, SOS, z  (difflib/tensor.py:10)


# attribute -> default, in the reference's order (model/config.py:9-45)
DEFAULTS = dict(
    args_dim=256, n_args=11, n_commands=N_COMMANDS, dropout=0.1, model_type="transformer", encode_stages=1,
    decode_stages=1, use_resnet=True, use_vae=True, pred_mode="one_shot", rel_targets=False, label_condition=False,
    n_labels=100, dim_label=64, self_match=False, n_layers=4, n_layers_decode=4, n_heads=8, dim_feedforward=512,
    d_model=256, dim_z=256, max_num_groups=8, max_seq_len=30)


class _DefaultConfig:
    #: per-variant overrides applied on top of DEFAULTS
    VARIANT = {}

    def __init__(self, **overrides):
        for key, value in {**DEFAULTS, **self.VARIANT, **overrides}.items():
            setattr(self, key, value)
        if "max_total_len" not in overrides:
            self.max_total_len = self.max_num_groups * self.max_seq_len
        if "num_groups_proposal" not in overrides:
            self.num_groups_proposal = self.max_num_groups

    def get_model_args(self):
        """Dataset field names fed positionally to forward: encoder (commands, args), decoder (commands, args)[, label]
        -- the `_grouped` variants for one-stage models (model/config.py:47-60)."""
        def pair(stages, rel):
            sfx = "_grouped" if stages <= 1 else ""
            return ["commands" + sfx, ("args_rel" if rel else "args") + sfx]
        fields = pair(self.encode_stages, False) + pair(self.decode_stages, self.rel_targets)   # model/config.py:47-56
        return fields + (["label"] if self.label_condition else [])


class OneStageOneShot(_DefaultConfig):          # model/config.py:83-89
    VARIANT = dict(encode_stages=1, decode_stages=1)


class Hierarchical(_DefaultConfig):             # model/config.py:92-98
    VARIANT = dict(encode_stages=2, decode_stages=2)


class HierarchicalSelfMatching(_DefaultConfig):  # model/config.py:101-108
    VARIANT = dict(encode_stages=2, decode_stages=2, self_match=True)


class SketchRNN(_DefaultConfig):                # model/config.py:63-71 (rejected by check_supported: LSTM)
    VARIANT = dict(model_type="lstm", pred_mode="autoregressive", rel_targets=True)


class Sketchformer(_DefaultConfig):             # model/config.py:74-80
    VARIANT = dict(pred_mode="autoregressive", rel_targets=True)


def check_supported(cfg):
    """The hot path covers model_type=transformer, pred_mode=one_shot, ordered assignment.
    Everything else raises at construction: there is no slow path to fall back to."""
    if getattr(cfg, "model_type", "transformer") != "transformer":
        raise NotImplementedError("deepsvg_b200: model_type='lstm' is outside the accelerated path")
    pm = getattr(cfg, "pred_mode", "one_shot")
    if pm not in ("one_shot", "autoregressive"):
        raise NotImplementedError("deepsvg_b200: unknown pred_mode %r" % (pm,))
    if pm == "autoregressive" and not (cfg.encode_stages == 1 and cfg.decode_stages == 1):
        raise NotImplementedError("deepsvg_b200: the autoregressive decoder is covered for the one-stage model (Sketchformer)")
    if getattr(cfg, "self_match", False) and not (cfg.encode_stages == 2 and cfg.decode_stages == 2):
        raise NotImplementedError("deepsvg_b200: self_match=True expects the two-stage model (model.py:385)")
    if getattr(cfg, "self_match", False) and (cfg.num_groups_proposal > 16 or cfg.max_num_groups > cfg.num_groups_proposal):
        raise NotImplementedError("deepsvg_b200: self_match needs max_num_groups <= num_groups_proposal <= 16")
    if cfg.encode_stages not in (1, 2) or cfg.decode_stages not in (1, 2) or cfg.encode_stages != cfg.decode_stages:
        raise NotImplementedError("deepsvg_b200: encode_stages and decode_stages must both be 1 or both be 2")
    if cfg.d_model % 128 != 0 or cfg.d_model > 512:
        raise NotImplementedError("deepsvg_b200: d_model must be 128, 256 or 512")
    hd = cfg.d_model // cfg.n_heads
    if cfg.d_model % cfg.n_heads != 0 or hd not in (16, 32, 64):
        raise NotImplementedError("deepsvg_b200: head_dim must be 16, 32 or 64")
    if cfg.dim_feedforward % 8 or cfg.dim_z % 8 or cfg.dim_label % 8:
        raise NotImplementedError("deepsvg_b200: dim_feedforward, dim_z, dim_label must be multiples of 8")
    if cfg.n_args != 11 or cfg.n_commands != N_COMMANDS:
        raise NotImplementedError("deepsvg_b200: n_args=11 / n_commands=7 are fixed by the SVG token vocabulary")
    if (2 * cfg.args_dim if getattr(cfg, "rel_targets", False) else cfg.args_dim + 1) > 512:
        raise NotImplementedError("deepsvg_b200: at most 512 classes per argument slot (args_dim <= 511, or 256 with rel_targets)")
    # sequence lengths the attention kernels hold on chip (csrc/attention.cu: the backward keeps Q, K, V, dO and two
    # L x L tiles of one (sequence, head) pair in shared memory; csrc/attention_mma.cu: L <= 80 on the tensor-core path)
    two = cfg.encode_stages == 2
    longest = (cfg.max_seq_len if two else cfg.max_total_len) + 2
    if two:
        longest = max(longest, cfg.max_num_groups, cfg.num_groups_proposal)
    if 4 * (4 * longest * hd + 2 * longest * (longest | 1) + 3) > 227 * 1024:
        raise NotImplementedError("deepsvg_b200: sequences of %d positions (head_dim %d) exceed the attention kernels' "
                                  "shared-memory tile (limits: 153 / 139 / 115 tokens for head_dim 16 / 32 / 64)"
                                  % (longest, hd))
