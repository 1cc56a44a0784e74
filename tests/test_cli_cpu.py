"""``python -m hefl_b200``: the command-line launcher of the product path, on CPU (single process,
checkpoint + resume, and the in-process simulation mode)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["--device", "cpu", "--model", "cnn2", "--image-size", "28", "--in-channels", "1", "--num-classes", "10",
        "--batch-size", "8", "--local-epochs", "1", "--steps-per-epoch", "2", "--val-steps", "1",
        "--he-preset", "n2048_l1", "--nn-backend", "cudnn", "--dtype", "fp32"]


def _run(args, env_extra=None):
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    if env_extra:
        env.update(env_extra)
    return subprocess.run([sys.executable, "-m", "hefl_b200", *args], capture_output=True, text=True, cwd=ROOT,
                          env=env, timeout=600)


def test_cli_runs_rounds_logs_and_resumes(tmp_path):
    ck, log = str(tmp_path / "ck.pt"), str(tmp_path / "run.jsonl")
    r = _run(BASE + ["--rounds", "2", "--checkpoint", ck, "--log-jsonl", log])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert [l["round"] for l in lines] == [0, 1] and all(l["loss"] == l["loss"] for l in lines)
    assert os.path.exists(ck) and len(open(log).read().splitlines()) == 2
    # resume: the third round only
    r = _run(BASE + ["--rounds", "3", "--checkpoint", ck, "--log-jsonl", log])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert [l["round"] for l in lines] == [2] and "resumed" in r.stderr


def test_cli_env_overrides_and_simulation_mode():
    r = _run(BASE + ["--simulate", "--rounds", "1"], env_extra={"HEFL_CLIENTS": "3"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["mode"] == "simulate" and d["clients"] == 3 and d["max_abs_err_vs_plaintext"] < 1e-4
