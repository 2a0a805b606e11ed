"""Constructor cases (option dicts as utils.model_init hands them over: all strings) shared by the
structure-golden generator and the CPU drop-in test.  Option names = proto/*.proto of the reference."""


def rec(prefix, lay, bn=True, bidir=True, act="relu", orth=True, drop=0.2):
    n = len(lay)
    return {
        f"{prefix}_lay": ",".join(map(str, lay)), f"{prefix}_drop": ",".join([str(drop)] * n),
        f"{prefix}_use_laynorm_inp": "False", f"{prefix}_use_batchnorm_inp": "False",
        f"{prefix}_use_laynorm": ",".join(["False"] * n), f"{prefix}_use_batchnorm": ",".join([str(bn)] * n),
        f"{prefix}_bidir": str(bidir), f"{prefix}_act": ",".join([act] * n), f"{prefix}_orthinit": str(orth),
        "use_cuda": "False", "to_do": "train",
        # sections also carry arch_* / opt_* keys (utils.py:2057 passes the whole section)
        "arch_name": "x", "arch_lr": "0.0004", "arch_opt": "rmsprop",
    }


def mlp(lay, drop, bn, ln, act, ln_inp=False, bn_inp=False):
    return {
        "dnn_lay": ",".join(map(str, lay)), "dnn_drop": ",".join(map(str, drop)),
        "dnn_use_laynorm_inp": str(ln_inp), "dnn_use_batchnorm_inp": str(bn_inp),
        "dnn_use_batchnorm": ",".join(map(str, bn)), "dnn_use_laynorm": ",".join(map(str, ln)),
        "dnn_act": ",".join(act), "use_cuda": "False", "to_do": "train", "arch_name": "y",
    }


def conv(prefix, n_filt, len_filt, pool, ln=True, ln_inp=False, act="relu", drop=0.15):
    n = len(n_filt)
    j = lambda v: ",".join(map(str, v))
    o = {f"{prefix}_N_filt": j(n_filt), f"{prefix}_len_filt": j(len_filt), f"{prefix}_max_pool_len": j(pool),
         f"{prefix}_use_laynorm_inp": str(ln_inp), f"{prefix}_use_batchnorm_inp": "False",
         f"{prefix}_use_laynorm": j([ln] * n), f"{prefix}_use_batchnorm": j([False] * n), f"{prefix}_act": j([act] * n),
         f"{prefix}_drop": j([drop] * n), "use_cuda": "False", "to_do": "train", "arch_name": "z"}
    if prefix == "sinc":
        o.update(sinc_sample_rate="16000", sinc_min_low_hz="50", sinc_min_band_hz="50")
    return o


CASES = {
    # the shapes of the shipped recipes (cfg/TIMIT_baselines/*.cfg, cfg/Librispeech_baselines/*.cfg), scaled down
    "ligru_timit": ("liGRU", rec("ligru", [55, 55, 55, 55, 55]), 40),
    "ligru_uni_nobn": ("liGRU", rec("ligru", [30, 20], bn=False, bidir=False, act="tanh", orth=False), 13),
    "gru_timit": ("GRU", rec("gru", [55, 55, 55, 55, 55], act="tanh"), 40),
    "lstm_timit": ("LSTM", rec("lstm", [55, 55, 55, 55], act="tanh"), 40),
    "rnn_timit": ("RNN", rec("rnn", [55, 55, 55, 55]), 40),
    "minimalgru": ("minimalGRU", rec("minimalgru", [48, 48], act="relu"), 23),
    "mlp_head": ("MLP", mlp([193], [0.0], [False], [False], ["softmax"]), 110),
    "mlp_timit": ("MLP", mlp([102, 102, 102, 102, 193], [0.15] * 4 + [0.0], [True] * 4 + [False], [False] * 5,
                              ["relu"] * 4 + ["softmax"]), 429),
    # cfg/TIMIT_baselines/TIMIT_SincNet_raw.cfg / TIMIT_CNN_fbank.cfg in miniature
    "sincnet_raw": ("SincNet", conv("sinc", [16, 12, 12, 12], [33, 5, 5, 3], [3, 3, 3, 2], ln=True, ln_inp=True), 800),
    "cnn_fbank": ("CNN", conv("cnn", [20, 12, 12], [10, 3, 3], [3, 2, 1]), 120),
    "mlp_ln_inp": ("MLP", mlp([20, 9], [0.1, 0.0], [False, False], [True, False], ["leaky_relu", "softmax"],
                               ln_inp=True, bn_inp=True), 11),
}
