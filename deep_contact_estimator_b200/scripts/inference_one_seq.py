#!/usr/bin/env python
"""`python -m deep_contact_estimator_b200.scripts.inference_one_seq --config_name <yaml>`

Same flow and config keys as /root/reference/src/inference_one_seq.py:137-176
(config/inference_one_seq_params.yaml), over the B200 path: the whole log is classified by one
`dce_stream` call.  Differences from the reference, all needed to run at all in a current
environment: `yaml.safe_load` (PyYAML >= 6), `torch.load` retried with `weights_only=False` for
checkpoints that hold numpy scalars (src/train.py:145-153), and an LCM log writer that does not
need liblcm (`lcm_wire.EventLog`) when `import lcm` fails.
"""
import argparse
import os
import time

import numpy as np
import torch
import yaml

from .. import contact_cnn, contact_dataset, inference, inference_and_compute_acc, decimal2binary, lcm_wire
from torch.utils.data import DataLoader


def load_checkpoint(path, map_location=None):
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        return torch.load(path, map_location=map_location, weights_only=False)


def save2mat(pred, config):
    """src/inference_one_seq.py:64-89 (needs scipy and the .mat with the raw signals)."""
    import scipy.io as sio
    ws = config["window_size"]
    mat_raw_data = sio.loadmat(config["mat_data_path"])
    data = np.load(config["data_path"])
    label = decimal2binary(torch.from_numpy(np.load(config["label_path"]))).reshape(-1, 4)
    out = {"contacts_est": pred.cpu().numpy(), "contacts_gt": label[ws - 1:, :].numpy(),
           "q": data[ws - 1:, :12], "qd": data[ws - 1:, 12:24], "imu_acc": data[ws - 1:, 24:27],
           "imu_omega": data[ws - 1:, 27:30], "p": data[ws - 1:, 30:42], "v": data[ws - 1:, 42:54],
           "control_time": mat_raw_data["control_time"].flatten().tolist()[ws - 1:],
           "imu_time": mat_raw_data["imu_time"].flatten().tolist()[ws - 1:],
           "tau_est": mat_raw_data["tau_est"][ws - 1:], "F": mat_raw_data["F"][ws - 1:]}
    sio.savemat(config["mat_save_path"], out)
    print("Saved data to mat!")


def save2lcm(pred, config, mat_data=None):
    """src/inference_one_seq.py:91-133: three events per step (leg_control_data, contact, microstrain)."""
    if mat_data is None:
        import scipy.io as sio
        mat_data = sio.loadmat(config["mat_data_path"])
    ws = config["window_size"]
    pred = pred.cpu().numpy()
    utime = int(time.time() * 10 ** 6)
    imu_time = np.asarray(mat_data["imu_time"]).flatten().tolist()
    with lcm_wire.EventLog(config["lcm_save_path"], mode="w", overwrite=True) as log:
        for idx, _ in enumerate(imu_time[ws - 1:]):
            di = idx + ws - 1
            t = utime + int(10 ** 6 * imu_time[di])
            log.write_event(t, "leg_control_data", lcm_wire.encode_leg_control_data(
                mat_data["q"][di], mat_data["qd"][di], mat_data["p"][di], mat_data["v"][di], mat_data["tau_est"][di]))
            log.write_event(t, "contact", lcm_wire.encode_contact(4, imu_time[di], pred[idx]))
            log.write_event(t, "microstrain", lcm_wire.encode_microstrain(
                mat_data["imu_quat"][di], mat_data["imu_rpy"][di], mat_data["imu_omega"][di], mat_data["imu_acc"][di]))
    print("Saved data to lcm!")


def main(argv=None):
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    print("Using ", device)
    parser = argparse.ArgumentParser(description="Test the contact network")
    parser.add_argument("--config_name", type=str,
                        default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "config", "inference_one_seq_params.yaml"))
    args = parser.parse_args(argv)
    with open(args.config_name) as f:
        config = yaml.safe_load(f)

    dataset = contact_dataset(data_path=config["data_path"], label_path=config["label_path"],
                              window_size=config["window_size"], device=device)
    dataloader = DataLoader(dataset=dataset, batch_size=config["batch_size"])
    model = contact_cnn()
    checkpoint = load_checkpoint(config["model_load_path"], map_location=device)
    model.load_state_dict(checkpoint["model_state_dict"])
    model = model.eval().to(device)

    if config["calculate_accuracy"]:
        pred, acc, acc_per_leg = inference_and_compute_acc(dataloader, model, device)
        print("Accuracy in terms of class: %.4f" % acc)
        for leg in range(4):
            print("Accuracy of leg %d is: %.4f" % (leg, acc_per_leg[leg]))
        print("Accuracy is: %.4f" % (np.sum(acc_per_leg) / 4.0))
    else:
        pred = inference(dataloader, model, device)
    if config.get("save_mat"):
        save2mat(pred, config)
    if config.get("save_lcm"):
        save2lcm(pred, config)
    return pred


if __name__ == "__main__":
    main()
