"""CLI option dict (`pdict`) defaults of the reference (vacmap:177-296), shared by the test harnesses."""


def default_option(mode="H", **over):
    skips = {"L": (59., 40., 0.1), "H": (40., 40., 0.2)}.get(mode, (30., 30., 0.5))
    opt = {"mode": mode, "c": 100, "eqx": False, "md": False, "cigar2cg": False, "copycomments": False, "H": False,
           "fakecigar": False, "Q": False, "debug": False, "shortcs": True, "rg-id": "1", "local_kmersize": 9,
           "local_skipcost": skips[0], "golbal_skipcost": skips[1], "maxdivergence": skips[2],
           "golbal_maxdiff": 50, "local_maxdiff": 30, "markunbalancetra": mode in ("L", "H"),
           "nodiscard": mode not in ("L", "H")}
    opt.update(over)
    return opt
