"""Env configuration: the 17 keys `continuousDoubleAuctionEnv.__init__` reads
(reference: gym_continuousDoubleAuction/envs/continuousDoubleAuction_env.py:35-53) with the
standalone fallbacks of the reference's config/env_defaults.json, plus the structural constants
of config/tunable_constants.json that shape the spaces (k_rows=10, book_rows=4, extra_dim=2,
category_n=9, price_offset_n=3).  If $CDA_CONFIG_DIR points at a config tree with the same JSON
files they override these values, as in the reference (config_loader.py:42).
"""
import json
import os

ENV_DEFAULTS = {
    "num_of_agents": 5, "init_cash": 1000000, "tick_size": 1, "tape_display_length": 10,
    "max_step": 64, "is_render": True, "n_hist": 4,
    "initial_price_min": 10, "initial_price_max": 100,
    "min_size": 1, "mkt_max_size": 100, "limit_size_multiple": 10,
    "order_penalty": 0.1, "trade_penalty": 0.05, "drawdown_penalty": 0.2,
    "passive_bonus": 0.1, "loss_multiplier": 1.5,
}
K_ROWS, BOOK_ROWS, EXTRA_DIM = 10, 4, 2
SNAPSHOT_DIM = K_ROWS * BOOK_ROWS + EXTRA_DIM
CATEGORY_N, PRICE_OFFSET_N = 9, 3
BOOK_ROW_ORDER = ("bid_price", "bid_size", "ask_price", "ask_size")


def env_defaults():
    d = dict(ENV_DEFAULTS)
    cfg_dir = os.environ.get("CDA_CONFIG_DIR")
    if cfg_dir:
        path = os.path.join(cfg_dir, "env_defaults.json")
        with open(path) as f:           # a missing file raises, like the reference's loader
            env = json.load(f)["environment"]
        d.update({k: v for k, v in env.items() if not k.startswith("_")})
    return d


def resolve(config):
    """Merge a (possibly partial) env config dict over the defaults; validate like the reference."""
    cfg = env_defaults()
    unknown = [k for k in (config or {}) if k not in cfg and k not in ("num_markets", "device", "order_capacity", "fill_capacity", "decimal_ledger")]
    cfg.update(config or {})
    if float(cfg["tick_size"]) != int(cfg["tick_size"]) or int(cfg["tick_size"]) < 1:
        raise ValueError("cda_b200 supports integral tick_size >= 1 only "
                         "(the reference does not quantise prices on other grids)")
    if float(cfg["init_cash"]) != int(cfg["init_cash"]) or int(cfg["init_cash"]) <= 0:
        raise ValueError("init_cash must be a positive integer")
    cfg["_unknown_keys"] = unknown
    return cfg
