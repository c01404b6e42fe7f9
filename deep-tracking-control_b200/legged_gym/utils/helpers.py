"""Checkpoint discovery and resume wiring of the training path: `get_load_path`
(legged_gym/utils/helpers.py:73-95) and the runner construction + resume of `TaskRegistry.make_alg_runner`
(legged_gym/utils/task_registry.py:100-128).  Scene construction, argument parsing and the task registry itself are out of scope;
the environment is whatever `LeggedRobotDTC` the caller built."""
import os
from datetime import datetime

from ..envs.base.cfg_resolve import class_to_dict  # noqa: F401


def get_load_path(root, load_run=-1, checkpoint=-1):
    """Path of `model_<checkpoint>.pt` inside run directory `load_run` under `root`; -1 = the last run (lexicographic order of the
    directory names, 'exported' ignored) / the highest-numbered model file.  Raises ValueError when `root` holds no run."""
    try:
        runs = sorted(os.listdir(root))
        if "exported" in runs:
            runs.remove("exported")
        last_run = os.path.join(root, runs[-1])
    except Exception:
        raise ValueError("No runs in this directory: " + root)
    load_run = last_run if load_run == -1 else os.path.join(root, load_run)
    if checkpoint == -1:
        models = [f for f in os.listdir(load_run) if "model" in f]
        models.sort(key=lambda m: "{0:0>15}".format(m))
        model = models[-1]
    else:
        model = "model_{}.pt".format(checkpoint)
    return os.path.join(load_run, model)


def make_alg_runner(env, train_cfg, log_root=None, device="cuda:0"):
    """OnPolicyRunner over `env` with the reference's log-directory naming and resume behaviour: when
    `train_cfg.runner.resume` is set, the checkpoint picked by `get_load_path(log_root, load_run, checkpoint)` is loaded (model,
    main optimizer state, iteration counter) BEFORE a new run directory is used.  Returns (runner, train_cfg)."""
    from ...rsl_rl.runners import OnPolicyRunner
    cfg = class_to_dict(train_cfg) if not isinstance(train_cfg, dict) else train_cfg
    rc = cfg["runner"]
    log_dir = None if log_root is None else os.path.join(log_root, datetime.now().strftime("%b%d_%H-%M-%S") + "_" + str(rc.get("run_name", "")))
    runner = OnPolicyRunner(env, cfg, log_dir, device=device)
    if rc.get("resume", False):
        if log_root is None:
            raise ValueError("resume needs the log root the previous run wrote to")
        resume_path = get_load_path(log_root, load_run=rc.get("load_run", -1), checkpoint=rc.get("checkpoint", -1))
        print(f"Loading model from: {resume_path}")
        runner.load(resume_path)
    return runner, train_cfg
