"""Minimal Keras-shaped façade over the PyTorch models so the reference's call sites keep
working unchanged: ``model.layers[i].get_weights()`` / ``set_weights`` (FLPyfhelin.py:205-221,
:271-278), ``model.get_weights()`` / ``set_weights`` (:151, :158), ``model.fit(...,
callbacks=[checkpoint, early, lr_red], epochs=)`` (:172, :193), ``model.save`` /
``load_model`` (:144, :175, :280), ``model.predict`` (notebook N:262), and data generators with
``flow_from_dataframe`` semantics (:59-70, :80-112).

Weights cross this boundary in Keras layouts (conv HWIO, dense [in,out]); files written by
``save`` are torch pickles whatever their extension says (``.hdf5`` in the reference).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ..config import FLConfig
from ..fl.data import BatchFeeder, SyntheticImageDataset, split_train_val
from ..fl.trainer import LocalTrainer
from ..models import ParamPack, create_model as _create_torch_model


# ---- callbacks (parameter holders; LocalTrainer.fit implements the behaviour) ----------------
class EarlyStopping:
    def __init__(self, monitor="loss", mode="min", patience=0, restore_best_weights=False, **_):
        self.monitor, self.patience, self.restore_best_weights = monitor, patience, restore_best_weights


class ReduceLROnPlateau:
    def __init__(self, monitor="loss", patience=10, verbose=0, factor=0.1, min_lr=0.0, **_):
        self.monitor, self.patience, self.factor, self.min_lr = monitor, patience, factor, min_lr


class ModelCheckpoint:
    def __init__(self, filepath, save_weights_only=False, save_best_only=False, verbose=0,
                 monitor="val_loss", mode="auto", **_):
        self.filepath, self.monitor = filepath, monitor


class History:
    def __init__(self, history: Dict[str, List[float]]):
        self.history = history


# ---- data ---------------------------------------------------------------------------------
class FrameIterator:
    """``ImageDataGenerator(...).flow_from_dataframe`` stand-in: loads the images listed in a
    DataFrame ['Path','Label'] (PIL, bilinear resize) into pinned uint8 memory once, then feeds
    batches. ``subset``/``validation_split`` follow Keras (the first fraction validates)."""

    def __init__(self, df, target_size, batch_size, shuffle, subset=None, validation_split=0.0,
                 augment=False, channels=3, device: Optional[torch.device] = None, seed: int = 0):
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.augment = augment
        labels = sorted(set(df["Label"]))
        self.class_indices = {l: i for i, l in enumerate(labels)}
        rows = list(zip(df["Path"], df["Label"]))
        n = len(rows)
        nval = int(n * validation_split)
        if subset == "validation":
            rows = rows[:nval]
        elif subset == "training":
            rows = rows[nval:]
        self.filenames = [r[0] for r in rows]
        self.classes = np.array([self.class_indices[r[1]] for r in rows], dtype=np.int64)
        self.samples = len(rows)
        print(f"Found {self.samples} validated image filenames belonging to {len(labels)} classes.")
        H, W = target_size
        imgs = torch.zeros(self.samples, H, W, channels, dtype=torch.uint8)
        for i, path in enumerate(self.filenames):
            imgs[i] = torch.from_numpy(_load_image(path, (H, W), channels))
        ds = SyntheticImageDataset.__new__(SyntheticImageDataset)
        pin = torch.cuda.is_available()
        ds.images = imgs.pin_memory() if pin else imgs
        lab = torch.from_numpy(self.classes)
        ds.labels = lab.pin_memory() if pin else lab
        ds.n, ds.classes = self.samples, len(labels)
        self.dataset = ds
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.feeder = BatchFeeder(ds, range(self.samples), batch_size, self.device, shuffle=shuffle, seed=seed)

    def __len__(self):
        return self.feeder.steps


def _load_image(path: str, size, channels: int) -> np.ndarray:
    if path.endswith(".npy"):
        arr = np.load(path)
    else:
        from PIL import Image

        img = Image.open(path).convert("RGB" if channels == 3 else "L").resize((size[1], size[0]), Image.BILINEAR)
        arr = np.asarray(img)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    if arr.shape[:2] != tuple(size):
        t = torch.from_numpy(arr).permute(2, 0, 1)[None].float()
        t = torch.nn.functional.interpolate(t, size=size, mode="bilinear", align_corners=False)
        arr = t[0].permute(1, 2, 0).round().clamp(0, 255).byte().numpy()
    return np.ascontiguousarray(arr[:, :, :channels]).astype(np.uint8)


# ---- model --------------------------------------------------------------------------------
class KLayer:
    def __init__(self, model: "KModel", index: int, kind: str, keys: Sequence[str]):
        self._m, self.index, self.kind, self._keys = model, index, kind, list(keys)
        self.name = f"{kind}_{index}"

    def get_weights(self) -> List[np.ndarray]:
        d = self._m.pack.to_keras_dict()
        return [d[k] for k in self._keys]

    def set_weights(self, weights: Sequence[np.ndarray]) -> None:
        d = self._m.pack.to_keras_dict()
        for k, w in zip(self._keys, weights):
            d[k] = np.asarray(w, dtype=np.float32).reshape(d[k].shape)
        self._m.pack.from_keras_dict(d)
        self._m._weights_changed()


class KModel:
    def __init__(self, cfg: FLConfig, device: Optional[torch.device] = None):
        self.cfg = cfg
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.net = _create_torch_model(cfg.model, cfg.in_channels, cfg.num_classes, cfg.image_size).to(self.device)
        self.pack = ParamPack(self.net)
        self._trainer: Optional[LocalTrainer] = None
        self.layers: List[KLayer] = []
        keys = self.pack.keras_order_keys()
        for li, (kind, mod) in enumerate(self.net.keras_layers()):
            mine = [k for k in keys if k.startswith(f"c_{li}_")]
            self.layers.append(KLayer(self, li, kind, mine))
        self.stop_training = False

    # trainer is created lazily so that pure weight shuffling never touches CUDA graphs
    def _get_trainer(self) -> LocalTrainer:
        if self._trainer is None:
            backend = self.cfg.nn_backend
            if backend == "tcgen05" and not (self.device.type == "cuda" and self.cfg.model == "medcnn"):
                backend = "cudnn"
            self._trainer = LocalTrainer(self.net, self.pack, self.cfg, self.device, backend=backend)
        return self._trainer

    def _weights_changed(self) -> None:
        if self._trainer is not None and self._trainer.engine is not None:
            self._trainer.engine.after_restore()

    def compile(self, **_):
        return None

    def get_weights(self) -> List[np.ndarray]:
        return list(self.pack.to_keras_dict().values())

    def set_weights(self, weights: Sequence[np.ndarray]) -> None:
        keys = self.pack.keras_order_keys()
        self.pack.from_keras_dict({k: np.asarray(w, dtype=np.float32) for k, w in zip(keys, weights)})
        self._weights_changed()

    def fit(self, train_ds: FrameIterator, validation_data: Optional[FrameIterator] = None,
            callbacks: Sequence = (), epochs: int = 1, verbose: int = 1) -> History:
        tr = self._get_trainer()
        tr.augment = bool(train_ds.augment)
        es = next((c for c in callbacks if isinstance(c, EarlyStopping)), None)
        rl = next((c for c in callbacks if isinstance(c, ReduceLROnPlateau)), None)
        ck = next((c for c in callbacks if isinstance(c, ModelCheckpoint)), None)

        def on_epoch(ep, st):
            if verbose:
                print(f"Epoch {ep + 1}/{epochs} - loss: {st.loss:.4f} - accuracy: {st.accuracy:.4f} - "
                      f"val_loss: {st.val_loss:.4f} - val_accuracy: {st.val_accuracy:.4f} - lr: {st.lr_scale * self.cfg.lr:.2e}")

        hist = tr.fit(train_ds.feeder, validation_data.feeder if validation_data is not None else None, epochs,
                      early_stopping=es.patience if es else None,
                      restore_best=es.restore_best_weights if es else False,
                      reduce_lr_patience=rl.patience if rl else None,
                      reduce_lr_factor=rl.factor if rl else 0.1, min_lr=rl.min_lr if rl else 0.0,
                      checkpoint_path=ck.filepath if ck else None, on_epoch=on_epoch)
        return History({"loss": [h.loss for h in hist], "accuracy": [h.accuracy for h in hist],
                        "val_loss": [h.val_loss for h in hist], "val_accuracy": [h.val_accuracy for h in hist]})

    def predict(self, ds: FrameIterator) -> np.ndarray:
        """Softmax probabilities in dataset order (``shuffle=False`` iterators; notebook N:262)."""
        self.net.eval()
        outs = []
        n = ds.samples
        with torch.no_grad():
            for i in range(0, n, ds.batch_size):
                x = ds.dataset.images[i:i + ds.batch_size].to(self.device)
                xf = x.permute(0, 3, 1, 2).float() * (1.0 / 255.0)
                outs.append(torch.softmax(self.net(xf).float(), dim=1).cpu())
        return torch.cat(outs).numpy()

    def load_weights(self, path: str) -> None:
        ck = torch.load(path, map_location="cpu", weights_only=False)
        self.pack.load_flat(ck["flat"])
        self._weights_changed()

    def save(self, path: str) -> None:
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        torch.save({"format": "hefl_b200.kmodel", "config": self.cfg.to_json(),
                    "flat": self.pack.flat.detach().cpu()}, path)

    def summary(self) -> None:
        for l in self.layers:
            print(l.index, l.kind, [w.shape for w in l.get_weights()])


HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def _load_keras_hdf5(path: str, device: Optional[torch.device]) -> "KModel":
    """A real Keras ``model.save('main_model.hdf5')`` file (FLPyfhelin.py:144, :175, :269): weights are read from
    ``model_weights/<layer>/<layer>/{kernel,bias}:0`` in layer order and installed through the same ``c_{i}_{j}``
    mapping the ciphertext dictionaries use (Keras layouts: conv HWIO, dense [in, out]). Needs ``h5py``, which is
    not part of this image: the function is exercised only where h5py exists."""
    try:
        import h5py  # type: ignore
    except ImportError as e:  # pragma: no cover - h5py is absent in the build image
        raise RuntimeError(
            f"{path} is a Keras HDF5 file; reading it needs the optional dependency h5py (not installed). "
            "Files written by this package (torch pickles under the .hdf5 name) load without it.") from e
    import numpy as np

    with h5py.File(path, "r") as f:  # pragma: no cover
        g = f["model_weights"] if "model_weights" in f else f
        names = [n.decode() if isinstance(n, bytes) else n for n in g.attrs.get("layer_names", list(g.keys()))]
        arrays = []
        for ln in names:
            wn = [n.decode() if isinstance(n, bytes) else n for n in g[ln].attrs.get("weight_names", [])]
            arrays.append([np.asarray(g[ln][w]) for w in wn])
    m = KModel(FLConfig(), device)  # pragma: no cover
    wl = [l for l in m.layers if l.get_weights()]
    src = [a for a in arrays if a]
    if len(src) != len(wl):
        raise ValueError(f"{path}: {len(src)} weighted layers, this architecture has {len(wl)}")
    for layer, ws in zip(wl, src):
        layer.set_weights(ws)
    return m


def load_model(path: str, device: Optional[torch.device] = None) -> KModel:
    import json

    with open(path, "rb") as fh:
        if fh.read(8) == HDF5_MAGIC:
            return _load_keras_hdf5(path, device)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    cfg = FLConfig(**json.loads(ck["config"]))
    m = KModel(cfg, device)
    m.pack.load_flat(ck["flat"])
    m._weights_changed()
    return m
