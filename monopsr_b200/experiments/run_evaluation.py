"""python -m monopsr_b200.experiments.run_evaluation --config_path ... --data_split val --device 0
(src/monopsr/experiments/run_evaluation.py:12-45: evaluates every checkpoint of the experiment as it appears)"""
import argparse
import os

from ..core import config_utils, experiment


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_path", type=str, default=os.path.join(os.getcwd(), "configs", "monopsr_model_000.yaml"),
                    help="Path to the pipeline config")
    ap.add_argument("--data_split", type=str, default="val", help="Data split for evaluation")
    ap.add_argument("--device", type=str, default="0", help="CUDA device id")
    ap.add_argument("--data_dir", type=str, default=None)
    ap.add_argument("--max_polls", type=int, default=None, help="stop after this many empty polls (default: never)")
    return ap.parse_args(argv)


def main(argv=None, **kw):
    args = parse_args(argv)
    os.environ["CUDA_VISIBLE_DEVICES"] = args.device
    config = config_utils.parse_yaml_config(args.config_path, data_dir=args.data_dir)
    config.dataset_config.data_split = args.data_split
    return experiment.evaluate(config, data_dir=args.data_dir, max_polls=args.max_polls, **kw)


if __name__ == "__main__":
    main()
