#!/usr/bin/env python
"""`python -m deep_contact_estimator_b200.scripts.test --config_name <yaml>`

Same flow and config keys as /root/reference/src/test.py:113-220 (config/test_params.yaml):
evaluate the test split, print accuracy / precision / Jaccard / confusion statistics.  The
forward + argmax + bits of the whole split is one `dce_stream` call and the metrics of src/test.py:19-70 come
from counters accumulated on the device (`dce_accuracy_counts`: per-leg 2x2 and 16x16 class confusion matrices),
so no per-window array crosses PCIe; `compute_metrics` below is the reference's sklearn formulation, kept as the
checker the tests compare those counters with.
"""
import argparse
import os

import numpy as np
import torch
import yaml
from torch.utils.data import DataLoader

from .. import contact_cnn, contact_dataset, compute_accuracy, evaluate
from .inference_one_seq import load_checkpoint


def compute_metrics(bin_pred_arr, bin_gt_arr, pred_arr, gt_arr):
    from sklearn.metrics import precision_score, jaccard_score, confusion_matrix
    legs = ("leg_rf", "leg_lf", "leg_rh", "leg_lh")
    out = {
        "precision_of_class": precision_score(gt_arr, pred_arr, average="weighted"),
        "precision_of_all_legs": precision_score(bin_gt_arr.flatten(), bin_pred_arr.flatten()),
        "precision_of_legs": [precision_score(bin_gt_arr[:, i], bin_pred_arr[:, i]) for i in range(4)],
        "jaccard_of_class": jaccard_score(gt_arr, pred_arr, average="weighted"),
        "jaccard_of_all_legs": jaccard_score(bin_gt_arr.flatten(), bin_pred_arr.flatten()),
        "jaccard_of_legs": [jaccard_score(bin_gt_arr[:, i], bin_pred_arr[:, i]) for i in range(4)],
    }
    cm = {leg: confusion_matrix(bin_gt_arr[:, i], bin_pred_arr[:, i], labels=[0, 1]) for i, leg in enumerate(legs)}
    cm["total"] = sum(cm[leg] for leg in legs)
    out["confusion_mat"] = cm
    out["fn_rate"] = {k: m[0, 1] / (m[0, 0] + m[0, 1]) for k, m in cm.items()}       # as src/test.py:34-38
    out["fp_rate"] = {k: m[1, 0] / (m[1, 0] + m[1, 1]) for k, m in cm.items()}       # as src/test.py:40-44
    return out


def main(argv=None):
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    print("Using ", device)
    parser = argparse.ArgumentParser(description="Test the contact network")
    parser.add_argument("--config_name", type=str,
                        default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "config", "test_params.yaml"))
    args = parser.parse_args(argv)
    with open(args.config_name) as f:
        config = yaml.safe_load(f)
    test_data = contact_dataset(data_path=config["data_folder"] + "test.npy", label_path=config["data_folder"] + "test_label.npy",
                                window_size=config["window_size"], device=device)
    loader = DataLoader(dataset=test_data, batch_size=config["batch_size"])
    model = contact_cnn()
    model.load_state_dict(load_checkpoint(config["model_load_path"], map_location=device)["model_state_dict"])
    model = model.eval().to(device)
    acc, acc_per_leg, m = evaluate(loader, model)
    print("Test accuracy in terms of class is: %.4f" % acc)
    for leg in range(4):
        print("Accuracy of leg %d is: %.4f" % (leg, acc_per_leg[leg]))
    print("Accuracy is: %.4f" % (np.sum(acc_per_leg) / 4.0))
    print("Precision of class: %.4f, of all legs: %.4f" % (m["precision_of_class"], m["precision_of_all_legs"]))
    print("Jaccard of class: %.4f, of all legs: %.4f" % (m["jaccard_of_class"], m["jaccard_of_all_legs"]))
    print("False negative rate (total): %.4f, false positive rate (total): %.4f" % (m["fn_rate"]["total"], m["fp_rate"]["total"]))
    return acc, acc_per_leg, m


if __name__ == "__main__":
    main()
