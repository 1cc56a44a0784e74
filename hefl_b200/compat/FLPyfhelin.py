"""Drop-in counterpart of the reference library ``FLPyfhelin.py`` (/root/reference/FLPyfhelin.py).

Same 19 function names, signatures, module globals and on-disk contract (SURVEY.md §2.1 C1-C21,
Appendix C): ``publickey.pickle``, ``privatekey.pickle``, ``main_model.hdf5``,
``agg_model.hdf5``, ``weights/weights{k}.npy``, ``weights/client_{k}.ckpt``,
``weights/client_{k}.pickle`` = ``{'key': Pyfhel, 'val': {'c_{layer}_{tensor}': ndarray[PyCtxt]}}``.
The notebook's cell-3 sequence (N:233-272) runs unchanged against this module.

What is different underneath: PyTorch instead of Keras, hefl_b200's BFV-fractional/CKKS kernels
instead of Pyfhel/SEAL, batched kernels instead of per-scalar Python loops. Known defects of the
reference are fixed and flagged (Q3 ``gen_rekey``, Q6/Q10 throw-away models, Q8 missing
``weights/`` directory); Q1 (clients share one model object) is reproduced only when
``COMPAT_SEQUENTIAL_CLIENTS`` is set.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np

from ..config import FLConfig
from ..fl.data import prep_df as _prep_df, shard_range
from .keras_like import (EarlyStopping, FrameIterator, KModel, ModelCheckpoint, ReduceLROnPlateau,  # noqa: F401
                         load_model)
from .pyfhel_shim import PyCtxt, Pyfhel, PyPtxt  # noqa: F401

# ---- module globals (FLPyfhelin.py:31-36) ------------------------------------------------------
INIT_LR = 1e-3
EPOCHS = 10
BS = 32
SCALE = 1
input_shape = (int(256 * SCALE), int(256 * SCALE), 3)
image_size = (int(256 * SCALE), int(256 * SCALE))

# ---- knobs the reference lacks -----------------------------------------------------------------
MODEL_NAME = "medcnn"
NUM_CLASSES = 2
COMPAT_SEQUENTIAL_CLIENTS = False      # True reproduces quirk Q1 (FLPyfhelin.py:180 vs :184-193)
KEYGEN_SEED = None                     # None = fresh entropy, like the reference (Q12)
DATA_SEED = 0


def configure(image_side: int = 256, channels: int = 3, model: str = "medcnn", num_classes: int = 2,
              batch_size: int = 32, scale: int = 1) -> None:
    """Change the globals the reference forces users to edit in place."""
    global input_shape, image_size, MODEL_NAME, NUM_CLASSES, BS, SCALE
    SCALE = scale
    input_shape = (int(image_side * scale), int(image_side * scale), channels)
    image_size = (int(image_side * scale), int(image_side * scale))
    MODEL_NAME, NUM_CLASSES, BS = model, num_classes, batch_size


def _cfg() -> FLConfig:
    return FLConfig(model=MODEL_NAME, image_size=image_size[0], in_channels=input_shape[2],
                    num_classes=NUM_CLASSES, batch_size=BS, lr=INIT_LR, lr_decay=INIT_LR / 10)


def _ensure_weights_dir() -> None:
    os.makedirs("weights", exist_ok=True)     # the reference never creates it (Q8)


# ---- L0 data (FLPyfhelin.py:38-114) -------------------------------------------------------------
def prep_df(folder, shuffle=True):
    return _prep_df(folder, shuffle=shuffle)


def get_test_data(df_test, test_path):
    return FrameIterator(df_test, image_size, BS, shuffle=False, channels=input_shape[2])


def get_train_data(df_train, train_path, index, num_client):
    start, end = shard_range(len(df_train.index), index, num_client)     # :75-78
    df = df_train[start:end]
    train = FrameIterator(df, image_size, BS, shuffle=True, subset="training", validation_split=0.1,
                          augment=True, channels=input_shape[2], seed=DATA_SEED + index)
    val = FrameIterator(df, image_size, BS, shuffle=True, subset="validation", validation_split=0.1,
                        augment=True, channels=input_shape[2], seed=DATA_SEED + index)
    return train, val


# ---- L1 model (FLPyfhelin.py:118-177) -----------------------------------------------------------
def create_model(load_model_path=None):
    if load_model_path:
        return load_model(load_model_path)       # the reference builds a model first and discards it (Q10)
    return KModel(_cfg())


def save_weights(model, ind):
    _ensure_weights_dir()
    weights = np.array(model.get_weights(), dtype="object")
    np.save("weights/weights" + ind + ".npy", weights, allow_pickle=True)
    return


def load_weights(ind):
    weights = np.load("weights/weights" + ind + ".npy", allow_pickle=True)
    model = create_model()
    model.set_weights(list(weights))
    return model


def train_server(train_ds, val_ds, epoch=10):
    """Centralised (non-federated) baseline trainer (FLPyfhelin.py:161-177; never called by the notebook)."""
    _ensure_weights_dir()
    model = create_model()
    early = EarlyStopping(monitor="loss", mode="min", patience=3)
    lr_red = ReduceLROnPlateau(monitor="loss", patience=2, verbose=1, factor=0.3, min_lr=0.000001)
    checkpoint_path = "weights/main.ckpt"
    checkpoint = ModelCheckpoint(filepath=checkpoint_path, save_weights_only=True, save_best_only=True,
                                 verbose=1, monitor="accuracy", mode="auto")
    model.fit(train_ds, validation_data=val_ds, epochs=epoch, callbacks=[early, lr_red, checkpoint])
    model.load_weights(checkpoint_path)
    save_weights(model, "main")
    model.save("main_model.hdf5")
    return


# ---- L5 orchestration ---------------------------------------------------------------------------
def train_clients(dataframe, train_path, num_clients, epoch=10):
    _ensure_weights_dir()
    model = create_model("main_model.hdf5")
    for i in range(num_clients):
        if i > 0 and not COMPAT_SEQUENTIAL_CLIENTS:
            model = create_model("main_model.hdf5")          # true FedAvg: every client starts from the global model
        train_ds, val_ds = get_train_data(dataframe, train_path, i, num_clients)
        early = EarlyStopping(monitor="loss", mode="min", patience=5, restore_best_weights=True)
        lr_red = ReduceLROnPlateau(monitor="loss", patience=2, verbose=1, factor=0.3, min_lr=0.000001)
        checkpoint_path = "weights/client_" + str(i + 1) + ".ckpt"
        checkpoint = ModelCheckpoint(filepath=checkpoint_path, save_weights_only=True, save_best_only=True,
                                     verbose=1, monitor="accuracy", mode="auto")
        model.fit(train_ds, validation_data=val_ds, callbacks=[checkpoint, early, lr_red], epochs=epoch)
        save_weights(model, str(i + 1))
    return


def encrypt_export_weights(indx):
    HE = get_pk()
    model = load_weights(str(indx + 1))
    start = time.time()
    encrypted_weights = {}
    for i in range(len(model.layers)):
        weights = model.layers[i].get_weights()
        if weights != []:
            for j in range(len(weights)):
                # one ciphertext per scalar, like the reference loop (:216-217), in one batched launch
                encrypted_weights["c_" + str(i) + "_" + str(j)] = HE.encryptFracBatch(weights[j])
    end = time.time()
    print("Time to encrypt weights:", end - start)
    filename = "weights/client_" + str(indx + 1) + ".pickle"
    export_weights(filename, encrypted_weights)
    return


def export_weights(filename, encrypted_weights):
    HE = get_pk()
    dic = {"key": HE, "val": encrypted_weights}
    start = time.time()
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(filename, "wb") as handle:
        pickle.dump(dic, handle, protocol=pickle.HIGHEST_PROTOCOL)
    end = time.time()
    print("Time to export weights to pickle:", end - start)
    return


def export_encrypted_clients_weights(num_client):
    start = time.time()
    for i in range(num_client):
        encrypt_export_weights(i)
        print("Weights exported: Client", i + 1)
    end = time.time()
    print("Total time to encrypt and export:", end - start)
    return


def get_sk():
    with open("privatekey.pickle", "rb") as handle:
        key = pickle.load(handle)
    HE = key["HE"]
    HE.from_bytes_context(key["con"])
    HE.from_bytes_publicKey(key["pk"])
    HE.from_bytes_secretKey(key["sk"])
    return HE


def decrypt_import_weights(filename):
    start = time.time()
    dec_weights = decrypt_weights(filename)
    end = time.time()
    print("Time to decrypt:", end - start)
    model = create_model("main_model.hdf5")
    for i in range(len(model.layers)):
        weights = model.layers[i].get_weights()
        if weights != []:
            weight = []
            for j in range(len(weights)):
                weight.append(dec_weights["c_" + str(i) + "_" + str(j)])
            model.layers[i].set_weights(weight)
    model.save("agg_model.hdf5")
    return model


def decrypt_weights(filename):
    HE = get_sk()
    enc_weights = import_encrypted_weights(filename)
    dec_weights = {}
    for key in enc_weights:
        dec_weights[key] = HE.decryptFracBatch(enc_weights[key])
    return dec_weights


def import_encrypted_weights(filename):
    start = time.time()
    with open(filename, "rb") as handle:
        dct = pickle.load(handle)
    cweights = dct["val"]
    HE2 = dct["key"]
    if HE2._ctx is None:                # the pickled Pyfhel is an empty shell: rehydrate the public part
        HE2 = get_pk()
    enc_weights = {}
    for key in cweights:
        arr = cweights[key]
        shape = arr.shape
        weight = arr.flatten()
        for l in np.arange(len(weight)):
            weight[l]._pyfhel = HE2     # re-attach the context (FLPyfhelin.py:320-321)
        enc_weights[key] = weight.reshape(shape)
    end = time.time()
    print("Time to import:", end - start)
    return enc_weights


# ---- L2 keys ------------------------------------------------------------------------------------
def gen_pk(s=128, m=2048):
    HE = Pyfhel()
    HE.contextGen(p=65537, sec=s, m=m)
    HE.keyGen(seed=KEYGEN_SEED)
    keys = {"HE": HE, "con": HE.to_bytes_context(), "pk": HE.to_bytes_publicKey()}
    with open("publickey.pickle", "wb") as handle:
        pickle.dump(keys, handle, protocol=pickle.HIGHEST_PROTOCOL)
    return HE


def get_pk():
    with open("publickey.pickle", "rb") as handle:
        key = pickle.load(handle)
    HE2 = key["HE"]
    HE2.from_bytes_context(key["con"])
    HE2.from_bytes_publicKey(key["pk"])
    return HE2


def gen_rekey():
    """Relinearisation keys. The reference body refers to an undefined ``HE`` (Q3); this is the
    evident intent: rehydrate the key holder and call ``relinKeyGen(bitCount=1, size=5)``."""
    HE = get_sk()
    relinKeySize = 5
    HE.relinKeyGen(bitCount=1, size=relinKeySize)
    return HE


def aggregate_encrypted_weights(num_client):
    """Key-less server aggregation: sum of the clients' ciphertexts, times 1/num_client
    (FLPyfhelin.py:366-390). Never touches the secret key."""
    dct_weights = {}
    denom = float(1 / num_client)
    start = time.time()
    HE = get_pk()
    for i in range(num_client):
        filename = "weights/client_" + str(i + 1) + ".pickle"
        enc_weights = import_encrypted_weights(filename)
        for key in enc_weights:
            if i == 0:
                dct_weights[key] = np.zeros_like(enc_weights[key], dtype=PyCtxt)   # object array of int 0 (:380)
            dct_weights[key] = enc_weights[key] + dct_weights[key]                # PyCtxt.__add__ per element (:381)
    for key in dct_weights:
        dct_weights[key] = dct_weights[key] * denom                                # PyCtxt * float (:385)
    end = time.time()
    print("Time to aggregate:", end - start)
    return dct_weights
