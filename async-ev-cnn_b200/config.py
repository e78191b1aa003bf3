"""Flag / YAML schema of the reference (src/scripts/config.py:24-148) on argparse + PyYAML
(configargparse is not needed): `-c cfg.yml` supplies defaults, command-line flags override them,
unknown keys are ignored (the reference uses parse_known_args, config.py:147)."""
import argparse
import os

import yaml

from .streams import parse_layers


def layers_dict(text):
    try:
        return parse_layers(text)
    except Exception:
        raise argparse.ArgumentTypeError("Format must be 'name1=h1,w1,i1,o1 name2=h2,w2,12,02 name3=i3,o3 name4=i4,o4 ...'")


def boolean(v):
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('-c', '--config', required=True, help='config file path')
    p.add_argument('--batch_size', type=int, default=1)
    p.add_argument('--reader_threads', type=int, default=4)
    p.add_argument('--input_data_dir', type=str, default=os.path.join(os.path.dirname(__file__), '../data/nmnist'))
    p.add_argument('--file_format', type=str, default='n-data')
    p.add_argument('--restore_net', type=str, default=None)
    p.add_argument('--network', type=str, default='YoloEventCuda',
                   help="'YoloEventCuda' (B200 backend); the reference's names are 'YoloEventNumpy', 'YoloFrameNumpy', 'YoloFrameTf'")
    p.add_argument('--frame_h', type=int, default=124)
    p.add_argument('--frame_w', type=int, default=124)
    p.add_argument('--example_h', type=int, default=124)
    p.add_argument('--example_w', type=int, default=124)
    p.add_argument('--leak', type=float, default=0.00015)
    p.add_argument('--frame_delay', type=int, default=50)
    p.add_argument('--yolo_cnn_layers', type=layers_dict, default=None)
    p.add_argument('--yolo_cnn_padding', type=str, default='VALID')
    p.add_argument('--yolo_num_cells_h', type=int, default=4)
    p.add_argument('--yolo_num_cells_w', type=int, default=4)
    p.add_argument('--yolo_num_bbox', type=int, default=2)
    p.add_argument('--batch_event_size', type=int, default=1)
    p.add_argument('--batch_event_usec', type=int, default=None)
    # B200 backend extensions (ignored by the reference because it parses known args only)
    p.add_argument('--n_streams', type=int, default=1, help='independent event streams advanced together')
    p.add_argument('--device', type=int, default=0)
    p.add_argument('--max_samples', type=int, default=None, help='stop after this many test samples')
    return p


def config(argv=None):
    parser = build_parser()
    pre, _ = parser.parse_known_args(argv)
    with open(pre.config) as f:
        cfg = yaml.safe_load(f) or {}
    known = {a.dest: a for a in parser._actions}
    defaults = {}
    for key, val in cfg.items():
        if key in known and val is not None:
            act = known[key]
            defaults[key] = act.type(str(val)) if act.type is not None and not isinstance(val, (int, float)) or \
                (act.type is layers_dict) else val
    parser.set_defaults(**defaults)
    args, _ = parser.parse_known_args(argv)
    return args
