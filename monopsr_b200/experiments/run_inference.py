"""python -m monopsr_b200.experiments.run_inference --checkpoint_name monopsr_model_000 --data_split test
--ckpt_num 80000 --device 0        (src/monopsr/experiments/run_inference.py:13-75: the config is read back from the
experiment's output folder, <data_dir>/outputs/<checkpoint_name>/<checkpoint_name>.yaml)"""
import argparse
import os

from ..core import config_utils, experiment


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint_name", type=str, default="monopsr_model_000", help="Checkpoint name must be specified as a str.")
    ap.add_argument("--data_split", type=str, default="val", help="Data split must be specified e.g. val or test")
    ap.add_argument("--ckpt_num", nargs="+", default="all", help="Checkpoint number ex. 80000")
    ap.add_argument("--det_2d_score_thr", type=float, nargs="+", default=[0.2, 0.2, 0.2], help="2D detection score threshold.")
    ap.add_argument("--device", type=str, default="0", help="CUDA device id")
    ap.add_argument("--data_dir", type=str, default=None)
    return ap.parse_args(argv)


def main(argv=None, **kw):
    args = parse_args(argv)
    data_dir = args.data_dir or os.path.join(os.getcwd(), "data")
    config = config_utils.parse_yaml_config(
        data_dir + "/outputs/" + args.checkpoint_name + "/" + args.checkpoint_name + ".yaml", data_dir=data_dir)
    config.dataset_config.mscnn_thr = list(args.det_2d_score_thr)
    os.environ["CUDA_VISIBLE_DEVICES"] = args.device
    return experiment.inference(config, args.data_split, args.ckpt_num, data_dir=data_dir, **kw)


if __name__ == "__main__":
    main()
