"""Entry point mirroring src/scripts/run_networks.py:15-59 for the B200 backend.

    python -m async_ev_cnn_b200.run_networks -c configs/efcn_event_cuda.yml

`network:` selects the model class by name exactly like run_networks.py:33; `YoloEventCuda` pairs with
CudaEventRunner (the reference's if/elif at :51-56).  Without a dataset on disk (input_data_dir
missing) a SyntheticReader serves seeded N-Caltech101-shaped recordings.
"""
import os

from .config import config
from .models import YoloEventCuda
from .runner import CudaEventRunner, SyntheticReader


def make_reader(args):
    if args.input_data_dir and os.path.isdir(args.input_data_dir):
        raise NotImplementedError("dataset readers (src/readers/*) are outside the hot path; export recordings to "
                                  ".npy [N,3] (y,x,ts) or use the synthetic reader")
    return SyntheticReader(args.example_h, args.example_w, n_samples=max(args.n_streams, args.max_samples or 4))


def main(argv=None):
    args = config(argv)
    reader = make_reader(args)
    classes = {"YoloEventCuda": YoloEventCuda}
    if args.network not in classes:
        raise SystemExit("network %r is not provided by the B200 backend (use YoloEventCuda; the reference's own "
                         "YoloEventNumpy/YoloFrameNumpy/YoloFrameTf live in the reference repo)" % args.network)
    num_classes = args.yolo_cnn_layers[list(args.yolo_cnn_layers)[-1]][-1] - 5 * args.yolo_num_bbox \
        if args.yolo_cnn_layers else reader.num_classes()
    network = classes[args.network](args.frame_h, args.frame_w, num_classes, args.yolo_cnn_layers, args.yolo_cnn_padding,
                                    args.yolo_num_cells_h, args.yolo_num_cells_w, args.yolo_num_bbox, 0.1, args.leak,
                                    args.restore_net, None, n_streams=args.n_streams, device=args.device,
                                    max_events_per_step=max(2048, args.batch_event_size))
    graph = network.build_graph(None)
    runner = CudaEventRunner(args, reader)
    outs, times = runner.run(graph)
    print("Mean fw time ({} runs): {}".format(len(times), sum(times) / max(1, len(times))))
    return outs


if __name__ == "__main__":
    main()
