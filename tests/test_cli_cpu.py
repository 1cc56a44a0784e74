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


def test_cli_trains_on_an_image_folder(tmp_path):
    """--data-dir: the reference's folder layout (<dir>/<label>/<file>), decoded by the threaded loader."""
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(0)
    for label in ("NORMAL", "PNEUMONIA"):
        d = tmp_path / "Train" / label
        d.mkdir(parents=True)
        for i in range(20):
            img = rng.integers(0, 120, (20, 24, 3), dtype=np.uint8)        # not the model's size: resized on load
            if label == "PNEUMONIA":
                img[:8, :8] += 100
            Image.fromarray(img).save(d / f"{i:03d}.png")
    args = ["--device", "cpu", "--model", "cnn2", "--image-size", "28", "--in-channels", "3", "--num-classes", "2",
            "--batch-size", "8", "--local-epochs", "2", "--he-preset", "n2048_l1", "--nn-backend", "cudnn",
            "--dtype", "fp32", "--rounds", "1", "--data-dir", str(tmp_path / "Train"), "--test-dir", str(tmp_path / "Train")]
    r = _run(args)
    assert r.returncode == 0, r.stderr[-2000:]
    recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert recs[0]["round"] == 0 and recs[0]["loss"] == recs[0]["loss"]
    m = recs[-1]
    assert m["test_images"] == 40 and all(0.0 <= m[k] <= 1.0 for k in ("precision", "recall", "f1", "accuracy"))


def test_image_folder_dataset_shards_like_the_reference(tmp_path):
    import numpy as np
    from PIL import Image
    from hefl_b200.fl.data import ImageFolderDataset
    for label in ("a", "b", "c"):
        d = tmp_path / label
        d.mkdir()
        for i in range(7):
            Image.fromarray(np.full((9, 9, 3), 10 * i, dtype=np.uint8)).save(d / f"{i}.png")
    full = ImageFolderDataset(str(tmp_path), image_size=8, shuffle_seed=3, pin=False)
    assert len(full) == 21 and full.classes == 3 and full.images.shape == (21, 8, 8, 3)
    parts = [ImageFolderDataset(str(tmp_path), image_size=8, index=i, num_clients=2, shuffle_seed=3, pin=False) for i in range(2)]
    assert [len(p) for p in parts] == [10, 10]                       # int(21 / 2) each, remainder dropped (F:75-78)
    assert parts[0].filenames + parts[1].filenames == full.filenames[:20]
    assert parts[0].class_indices == parts[1].class_indices == {"a": 0, "b": 1, "c": 2}


def test_classification_metrics_match_sklearn_weighted():
    import numpy as np
    import torch
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score
    from hefl_b200.fl import classification_metrics
    rng = np.random.default_rng(1)
    for k in (2, 3, 5):
        yt = rng.integers(0, k, 200); yp = rng.integers(0, k, 200)
        yp[:40] = yt[:40]
        if k == 5:
            yp[yp == 4] = 0                                   # a class that is never predicted (zero_division path)
        m = classification_metrics(torch.from_numpy(yt), torch.from_numpy(yp), k)
        assert abs(m["precision"] - precision_score(yt, yp, average="weighted", zero_division=0)) < 1e-9
        assert abs(m["recall"] - recall_score(yt, yp, average="weighted", zero_division=0)) < 1e-9
        assert abs(m["f1"] - f1_score(yt, yp, average="weighted", zero_division=0)) < 1e-9
        assert abs(m["accuracy"] - accuracy_score(yt, yp)) < 1e-12
