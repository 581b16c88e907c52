"""python -m monopsr_b200.experiments.run_training --config_path configs/monopsr_model_000.yaml --data_split train
--device 0            (data parallel: torchrun --nproc-per-node N -m monopsr_b200.experiments.run_training ...;
same flags as src/monopsr/experiments/run_training.py:19-44; the config copy / backup in the
experiment's output folder is :49-66)"""
import argparse
import datetime
import filecmp
import os
import shutil

from ..core import config_utils, experiment


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_path", type=str, default=os.path.join(os.getcwd(), "configs", "monopsr_model_000.yaml"),
                    help="Path to the config")
    ap.add_argument("--data_split", type=str, default="train", help="Data split for training")
    ap.add_argument("--device", type=str, default="0", help="CUDA device id")
    ap.add_argument("--data_dir", type=str, default=None, help="root of outputs/ and detections/ (default ./data)")
    ap.add_argument("--pretrained_checkpoint", type=str, default=None,
                    help="object-detection-API ResNet-101 checkpoint prefix for both encoders")
    return ap.parse_args(argv)


def keep_config_copy(config_path, config, log=print):
    os.makedirs(config.exp_output_dir, exist_ok=True)
    copy_path = config.exp_output_dir + "/{}.yaml".format(config.config_name)
    if not os.path.exists(copy_path):
        shutil.copyfile(config_path, copy_path)
    elif not filecmp.cmp(config_path, copy_path):
        stamp = str(datetime.datetime.now())
        shutil.copyfile(copy_path, copy_path + "." + stamp)
        shutil.copyfile(config_path, copy_path)
        log("Config file has changed since ", stamp)
    return copy_path


def main(argv=None, **kw):
    args = parse_args(argv)
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:          # under torchrun every rank takes cuda:LOCAL_RANK instead
        os.environ["CUDA_VISIBLE_DEVICES"] = args.device
    config = config_utils.parse_yaml_config(args.config_path, data_dir=args.data_dir)
    keep_config_copy(args.config_path, config)
    config.dataset_config.data_split = args.data_split
    return experiment.train(config, data_dir=args.data_dir, pretrained_checkpoint=args.pretrained_checkpoint, **kw)


if __name__ == "__main__":
    main()
