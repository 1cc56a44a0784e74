#!/usr/bin/env python
"""Script form of the reference driver notebook ``Encrypted FL Main-Rel.ipynb`` (cells 0-6,
raw lines N:10-32, N:53-65, N:87-95, N:226-274, N:342-347, N:400-405, N:415-429).

    python examples/encrypted_fl_main.py --workdir /tmp/run --synthetic 28 --clients 2 --epochs 1

``--synthetic S`` writes a tiny synthetic image folder tree (image/Train/<label>/*.png,
image/Test/...) of SxS images so the whole pipeline runs offline; without it the script expects
the reference's ``image/Train`` and ``image/Test`` folders in the working directory.
"""
import argparse
import os
import pickle
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

# ---- cell 0 (N:10-32): imports and settings ----------------------------------------------------
import FLPyfhelin as FL  # noqa: E402
from FLPyfhelin import *  # noqa: E402,F401,F403
from Pyfhel import PyCtxt, Pyfhel, PyPtxt  # noqa: E402,F401


def make_synthetic_folders(root: str, side: int, n_train: int, n_test: int, seed: int = 0) -> None:
    from PIL import Image

    rng = np.random.default_rng(seed)
    for split, n in (("Train", n_train), ("Test", n_test)):
        for label in ("NORMAL", "PNEUMONIA"):
            d = os.path.join(root, "image", split, label)
            os.makedirs(d, exist_ok=True)
            for i in range(n // 2):
                img = rng.integers(0, 140, (side, side, 3), dtype=np.uint8)
                if label == "PNEUMONIA":
                    q = max(2, side // 3)
                    img[:q, :q, :] += 100
                Image.fromarray(img).save(os.path.join(d, f"{label}_{i:04d}.png"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workdir", default=".")
    ap.add_argument("--synthetic", type=int, default=0, help="side of synthetic images to generate (0 = use existing folders)")
    ap.add_argument("--clients", type=int, nargs="+", default=[2])     # num_of_client_list (N:233)
    ap.add_argument("--epochs", type=int, default=10)                  # epoch (N:30)
    ap.add_argument("--m", type=int, default=1024)                     # gen_pk(s=128, m=1024) (N:54)
    ap.add_argument("--model", default=None)
    ap.add_argument("--train-images", type=int, default=64)
    ap.add_argument("--test-images", type=int, default=32)
    args = ap.parse_args()
    os.makedirs(args.workdir, exist_ok=True)
    os.chdir(args.workdir)
    if args.synthetic:
        FL.configure(image_side=args.synthetic, channels=3, model=args.model or "cnn2", batch_size=8)
        if not os.path.isdir("image/Train"):
            make_synthetic_folders(".", args.synthetic, args.train_images, args.test_images)
    elif args.model:
        FL.configure(model=args.model)
    train_path, test_path, epoch = "image/Train", "image/Test", args.epochs

    # ---- cell 1 (N:53-65): key generation, private key file -----------------------------------
    HE = gen_pk(s=128, m=args.m)
    keys = {"HE": HE, "con": HE.to_bytes_context(), "pk": HE.to_bytes_publicKey(), "sk": HE.to_bytes_secretKey()}
    with open("privatekey.pickle", "wb") as handle:
        pickle.dump(keys, handle, protocol=pickle.HIGHEST_PROTOCOL)
    print(HE)

    # ---- cell 2 (N:87-95): inline get_sk (the notebook's copy failed with a NameError) ----------
    HE_sk = get_sk()
    print(HE_sk)

    # ---- cell 3 (N:226-274): the federated round -------------------------------------------------
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score

    p, r, f, a, t = [], [], [], [], []
    for num_client in args.clients:
        start = time.time()
        print("Prepare data")
        df_train = prep_df(train_path, shuffle=True)
        df_test = prep_df(test_path, shuffle=False)
        test_ds = get_test_data(df_test, test_path)          # the notebook passes train_path here (Q7)
        print("Create and save the global model")
        model = create_model()
        model.save("main_model.hdf5")
        print("Train clients")
        train_clients(df_train, train_path, num_client, epoch)
        print("Encrypt and export client weights")
        export_encrypted_clients_weights(num_client)
        print("Aggregate encrypted weights")
        main_model_dict = aggregate_encrypted_weights(num_client)
        filename = "weights/aggregated.pickle"
        export_weights(filename, main_model_dict)
        print("Decrypt the aggregated weights")
        agg_model = decrypt_import_weights(filename)
        preds = agg_model.predict(test_ds)
        predictions = [int(np.argmax(x)) for x in preds]
        p.append(precision_score(test_ds.classes, predictions, average="weighted", zero_division=0))
        r.append(recall_score(test_ds.classes, predictions, average="weighted", zero_division=0))
        f.append(f1_score(test_ds.classes, predictions, average="weighted", zero_division=0))
        a.append(accuracy_score(test_ds.classes, predictions))
        t.append(time.time() - start)

    # ---- cell 4 (N:342-347) and cell 5 (N:400-405): result tables --------------------------------
    metrics = pd.DataFrame([p, r, f, a], index=["precision", "recall", "f1", "accuracy"],
                           columns=[str(c) for c in args.clients])
    times = pd.DataFrame([t], index=["time"], columns=[str(c) for c in args.clients])
    print(metrics)
    print(times)

    # ---- cell 6 (N:415-429, unexecuted in the reference): plaintext export for comparison ---------
    model = load_weights("1")
    plain = {}
    for i in range(len(model.layers)):
        ws = model.layers[i].get_weights()
        for j, w in enumerate(ws):
            plain["c_" + str(i) + "_" + str(j)] = w
    export_weights("plainweights.pickle", plain)
    enc_size = os.path.getsize("weights/client_1.pickle")
    plain_size = os.path.getsize("plainweights.pickle")
    print(f"encrypted client file: {enc_size / 1e6:.2f} MB, plaintext file: {plain_size / 1e6:.3f} MB, "
          f"expansion x{enc_size / max(plain_size, 1):.0f}")
    return metrics, times


if __name__ == "__main__":
    main()
