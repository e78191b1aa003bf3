"""TEST INFRASTRUCTURE ONLY - CPU restatement of the steps either side of the hot path:

  f3  decode of the N-MNIST / N-Caltech101 5-byte event records     src/readers/file_reader.py:30-58
  f1  the runner's per-sample transform (zero-based ts, centre crop)  src/libs/runner.py:24-33, src/libs/utils.py:4-35
  f2  YOLO decode of the [h_cells, w_cells, C + 5B] head               src/libs/viz.py:27-46,125-148

Parity: `center_crop` and `convert_bboxes` are checked against the reference's own functions when
/root/reference is present (tests/test_frontend.py); `file_reader.py` cannot be imported here (it needs
`bitstring`), so the record decoder is pinned by hand-made known-answer records and by a round trip through a
restatement of the reference's own encoder (file_reader.py:60-75).
Only tests/ may import this module; the product path never does.
"""
import numpy as np


def encode_ndata(x, y, ts, p):
    """5 bytes per event, big-endian 40 bits: x[8] y[8] p[1] ts[23]   (file_reader.py:60-75)."""
    v = (np.asarray(x, np.uint64) << np.uint64(32)) + (np.asarray(y, np.uint64) << np.uint64(24)) + \
        (np.asarray(p, np.uint64) << np.uint64(23)) + np.asarray(ts, np.uint64)
    out = np.empty((v.size, 5), np.uint8)
    for i in range(5):
        out[:, i] = (v >> np.uint64(8 * (4 - i))) & np.uint64(0xff)
    return out.reshape(-1)


def read_ndata(raw):
    """raw uint8 [5n] -> (length, x, y, ts, p) int32, overflow records (y == 240) applied and removed.
    file_reader.py:36-58."""
    raw = np.uint32(np.asarray(raw, np.uint8))
    all_y = raw[1::5]
    all_x = raw[0::5]
    all_p = (raw[2::5] & 128) >> 7
    all_ts = ((raw[2::5] & 127) << 16) | (raw[3::5] << 8) | (raw[4::5])
    time_increment = 2 ** 13
    for overflow_index in np.where(all_y == 240)[0]:            # file_reader.py:45-48
        all_ts[overflow_index:] += time_increment
    td = np.where(all_y != 240)[0]
    x = np.array(all_x[td], dtype=np.int32)
    y = np.array(all_y[td], dtype=np.int32)
    ts = np.array(all_ts[td], dtype=np.int32)
    p = np.array(all_p[td], dtype=np.int32)
    return len(x), x, y, ts, p


def center_crop_events(x, y, ts, p, new_shape):
    """Event half of utils.center_crop (utils.py:4-28), including its naming quirk: the top margin is derived
    from the x extent and the left margin from the y extent."""
    new_h, new_w = new_shape
    new_top = (x.max() - x.min() - new_w) // 2
    new_left = (y.max() - y.min() - new_h) // 2
    inside = np.logical_and.reduce([x >= new_left, x < new_left + new_w, y >= new_top, y < new_top + new_h])
    nx, ny, nts, npol = x[inside].copy(), y[inside].copy(), ts[inside], p[inside]
    if nx.size:                       # the reference raises on an empty crop (min of an empty array); callers skip such samples
        nx -= nx.min()
        ny -= ny.min()
    return nx, ny, nts, npol


def data_transform(x, y, ts, p, example_shape, frame_shape):
    """runner.py:24-33: ts zero-based on the first event, centre crop when the frame is smaller than the
    recording, events stacked as (y, x, ts)."""
    if len(ts):
        ts = ts - ts[0]
    if tuple(example_shape) != tuple(frame_shape) and len(x):
        x, y, ts, p = center_crop_events(x, y, ts, p, frame_shape)
    return np.stack([y, x, ts], axis=-1).astype(np.int32), np.asarray(p, np.int32)


def convert_bboxes(bboxes, grid_h, grid_w, h_image, w_image, sqrt):
    """viz.py:27-46.  bboxes float32 [n, grid_h, grid_w, B, 4] cell-relative -> pixels (x, y, w, h)."""
    cell_idx_h = np.arange(grid_h, dtype=np.float32)
    cell_idx_w = np.arange(grid_w, dtype=np.float32)
    col_idx = np.reshape(np.tile(cell_idx_w, [grid_h]), [grid_h, grid_w])
    row_idx = np.tile(np.expand_dims(cell_idx_h, axis=-1), [1, grid_w])
    col_idx = np.reshape(col_idx, [1, grid_h, grid_w, *([1] * (bboxes.ndim - 3))])
    row_idx = np.reshape(row_idx, [1, grid_h, grid_w, *([1] * (bboxes.ndim - 3))])
    true_x = ((bboxes[..., 0:1] + col_idx) / grid_w) * w_image
    true_y = ((bboxes[..., 1:2] + row_idx) / grid_h) * h_image
    true_w = (np.square(bboxes[..., 2:3]) if sqrt else bboxes[..., 2:3]) * w_image
    true_h = (np.square(bboxes[..., 3:4]) if sqrt else bboxes[..., 3:4]) * h_image
    return np.concatenate([true_x, true_y, true_w, true_h], axis=-1)


def decode_head(net_predictions, h_grid, w_grid, num_classes, h_image, w_image, conf_threshold):
    """The decode half of viz.draw_bboxes (viz.py:131-148,165) for a batch of heads [n, h_grid, w_grid, C + 5B]:
    returns boxes [n, cells*B, 4] (x, y, w, h in pixels), conf [n, cells*B], valid = conf > threshold, and the
    label index argmax_c(class_c * conf) per box."""
    net_predictions = np.asarray(net_predictions, np.float32)
    n = net_predictions.shape[0]
    pred_label = net_predictions[..., :num_classes]
    pred_bbox = np.reshape(net_predictions[..., num_classes:], [n, h_grid, w_grid, -1, 5])
    pred_bbox_params = pred_bbox[..., 0:4]
    pred_bbox_conf = pred_bbox[..., 4:5]
    trans = convert_bboxes(pred_bbox_params, h_grid, w_grid, h_image, w_image, sqrt=True)
    pred_label = np.expand_dims(pred_label, axis=-2) * pred_bbox_conf
    trans = np.reshape(trans, [n, -1, 4])
    pred_label = np.reshape(pred_label, [n, -1, num_classes])
    conf = np.reshape(np.max(pred_bbox_conf, axis=-1), [n, -1])
    valid = conf > conf_threshold
    return trans.astype(np.float32), conf.astype(np.float32), valid, np.argmax(pred_label, axis=-1).astype(np.int32)
