"""Training driver: restates dyn_mult_view/multi_view_model/train.py (flags :18-22, conf
loading :40-60, loop :117-151, checkpoint cadence :25-31,95-103,134-136) for this build, plus
the CUDA-graph-captured step that the benchmark uses.

    python -m dynamic_multiview_3d_b200.train --hyper path/to/conf.py [--pretrained model123]
"""
import argparse
import importlib.util
import os
import re
import sys
import time
import types

import torch

SUMMARY_INTERVAL = 400   # train.py:25
VAL_INTERVAL = 500       # train.py:28
SAVE_INTERVAL = 10000    # train.py:31


class GraphedTrainStep:
    """model.train_step captured once into a CUDA graph and replayed: removes the per-launch
    host cost of the ~150 kernels of a step.  Inputs are copied into static device buffers
    (from pinned host memory or device tensors) on the same stream before each replay."""

    def __init__(self, model, warmup=3):
        self.model = model
        dev = model.device
        B, (H, W, Cc), V = model.batch_size, model.image_shape, model.viewpoint_dim
        self.image0 = torch.zeros((B, H, W, Cc), dtype=torch.float32, device=dev)
        self.image1 = torch.zeros((B, H, W, Cc), dtype=torch.float32, device=dev)
        self.disp = torch.zeros((B, V), dtype=torch.float32, device=dev)
        self.graph = None
        self.loss = None
        self.warmup = warmup
        self.launches_per_step = 0
        self._stage = None
        self._pending = False
        self._u8 = None

    def capture(self):
        from . import _lib
        m = self.model
        side = torch.cuda.Stream(device=m.device)
        side.wait_stream(torch.cuda.current_stream(m.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                m.train_step(self.image0, self.image1, self.disp)
        torch.cuda.current_stream(m.device).wait_stream(side)
        torch.cuda.synchronize(m.device)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = m.train_step(self.image0, self.image1, self.disp)
        self.launches_per_step = _lib.launch_count() - n0
        return self

    def _copy_in(self, image0, image1, disp):
        """Bring a batch into the graph's static fp32 buffers.  uint8 images (the reference's TFRecord pixel format,
        a quarter of the PCIe bytes) are converted on the device: float32(pixel) / 255 (read_tf_records.py:111)."""
        from . import _lib
        st = torch.cuda.current_stream(self.model.device).cuda_stream
        for dst, src in ((self.image0, image0), (self.image1, image1)):
            if src.dtype == torch.uint8:
                if not src.is_cuda:
                    if self._u8 is None:
                        self._u8 = {}
                    buf = self._u8.get(id(dst))
                    if buf is None:
                        buf = self._u8[id(dst)] = torch.empty(dst.shape, dtype=torch.uint8, device=dst.device)
                    buf.copy_(src, non_blocking=True)
                    src = buf
                _lib.call("dmv_u8_to_f32", src.data_ptr(), dst.data_ptr(), dst.numel(), 255.0, st)
            else:
                dst.copy_(src, non_blocking=True)
        self.disp.copy_(disp, non_blocking=True)

    def replay(self):
        """One more step on the batch that already sits in the static buffers (no copy at all)."""
        self.graph.replay()
        return self.loss

    def prefetch(self, image0, image1, disp):
        """Start the host->device copy of the NEXT step's batch on a copy stream into staging buffers; it overlaps
        the step that is replayed meanwhile.  The following __call__ with ``staged=True`` consumes it."""
        dev = self.model.device
        if self._stage is None:
            self._stage = [torch.empty(self.image0.shape, dtype=image0.dtype, device=dev), torch.empty(self.image1.shape, dtype=image1.dtype, device=dev),
                           torch.empty_like(self.disp)]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged_ev = torch.cuda.Event()
            self._consumed_ev = torch.cuda.Event()
            self._consumed_ev.record(torch.cuda.current_stream(dev))
        self._copy_stream.wait_event(self._consumed_ev)          # the previous staged batch has been moved out
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._stage, (image0, image1, disp)):
                dst.copy_(src, non_blocking=True)
            self._staged_ev.record(self._copy_stream)
        self._pending = True

    def __call__(self, image0=None, image1=None, disp=None, staged=False, prefetch_next=None):
        """One train step.  ``staged=True``: take the batch handed to prefetch() (a 25 us device-side move instead
        of a PCIe copy in front of the step).  ``prefetch_next=(image0, image1, disp)``: start copying the next
        step's batch before this step is replayed, so the PCIe transfer overlaps it."""
        if self.graph is None:
            if staged:
                torch.cuda.current_stream(self.model.device).wait_event(self._staged_ev)
                image0, image1, disp = self._stage
            self._copy_in(image0, image1, disp)
            self.capture()
        if staged:
            main = torch.cuda.current_stream(self.model.device)
            main.wait_event(self._staged_ev)
            self._copy_in(*self._stage)
            self._consumed_ev.record(main)
            self._pending = False
        else:
            self._copy_in(image0, image1, disp)
            if self._stage is not None:
                self._consumed_ev.record(torch.cuda.current_stream(self.model.device))
        if prefetch_next is not None:
            self.prefetch(*prefetch_next)
        self.graph.replay()
        return self.loss


def install_reference_aliases():
    """Let an unmodified reference conf.py import its model classes by the reference's
    module names (e.g. ``from highdim_angle import AppFlowHighDimAngle``, ``import dyn_mult_view``)."""
    from . import appearance_flow_model as afm
    names = {
        "appearance_flow_model": afm, "highdim_angle": afm, "lowdim_angle": afm, "appearance_flow_tinghui": afm,
    }
    try:
        from . import main_model as mm
        from . import multiobject_appflow as mo
        names["main_model"] = mm
        names["multiobject_appflow"] = mo
    except ImportError:
        pass
    for k, mod in names.items():
        sys.modules.setdefault(k, mod)
    if "dyn_mult_view" not in sys.modules:
        pkg = types.ModuleType("dyn_mult_view")
        pkg.__file__ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dyn_mult_view", "__init__.py")
        sys.modules["dyn_mult_view"] = pkg


def load_conf(path):
    """train.py:40-45: imp.load_source('hyperparams', conf_file).configuration"""
    if not os.path.exists(path):
        sys.exit("Experiment configuration not found")
    install_reference_aliases()
    spec = importlib.util.spec_from_file_location("hyperparams", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.configuration


def checkpoint_iteration(path):
    """train.py:95-103: the iteration number is recovered from the file name ``model<itr>``."""
    m = re.match(r".*model(\d+)", os.path.basename(path) if "/" in path else path)
    return int(m.group(1)) if m else 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--hyper", default="conf.py", help="hyperparameters configuration file")
    ap.add_argument("--visualize", default="", help="model within hyperparameter folder to visualise")
    ap.add_argument("--device", default="0", help="the value for CUDA_VISIBLE_DEVICES")
    ap.add_argument("--pretrained", default=None, help="path to model file from which to resume training")
    ap.add_argument("--synthetic", action="store_true", default=True, help="synthetic car-render batches (no TFRecords)")
    ap.add_argument("--tfrecords", action="store_true", help="read conf['data_dir'] TFRecords (read_tf_records.py) instead of synthetic batches")
    ap.add_argument("--num_iterations", type=int, default=None)
    args = ap.parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", str(args.device))
    conf = load_conf(args.hyper)
    if args.num_iterations is not None:
        conf["num_iterations"] = args.num_iterations
    from .appearance_flow_model import AppearanceFlowModel
    from .synthetic import make_batch
    from .main_model import Base_Prediction_Model
    Model = conf.get("model", Base_Prediction_Model if ("use_color" in conf or "use_depth" in conf) else AppearanceFlowModel)
    model = Model(conf, load_tfrec=True, build_loss=not args.visualize)
    out_dir = conf.get("output_dir", ".")
    os.makedirs(out_dir, exist_ok=True)
    itr_0 = 0
    if args.pretrained:
        model.load_state_dict(torch.load(args.pretrained, map_location="cpu"))
        itr_0 = checkpoint_iteration(args.pretrained) + 1
    V = "onehot19" if model.viewpoint_dim == 19 else "disp2"
    if args.visualize:                       # train.py:60-65, 86-93: restore <output_dir>/<visualize> and write the figures
        conf["visualize"] = args.visualize
        ck = os.path.join(out_dir, args.visualize)
        if os.path.exists(ck):
            model.load_state_dict(torch.load(ck, map_location="cpu"))
        b = make_batch(model.batch_size, model.image_shape[0], V, seed=99)
        info = model.visualize(*(torch.from_numpy(b[k]).to(model.device) for k in ("image0", "image1", "disp")))
        print("loss", info["loss"])
        print("max resample coord:", info["max_resample_coord"])
        return
    step = GraphedTrainStep(model)
    t_iter = []
    reader = None
    if args.tfrecords:                       # train.py:57 load_tfrec=True -> read_tf_records.build_tfrecord_input
        from .read_tf_records import build_tfrecord_input
        reader = build_tfrecord_input(dict(conf, image_size=model.image_shape[0]), training=True)
    for itr in range(itr_0, conf["num_iterations"] + 1):
        t0 = time.time()
        if reader is not None:               # uint8 pixels go over PCIe as stored; /255 happens on the device (_copy_in)
            fb = reader.next_batch()
            batch = {"image0": fb["image0"], "image1": fb["image1"], "disp": fb["displacement"]}
        else:
            batch = make_batch(model.batch_size, model.image_shape[0], V, seed=1234 + itr)
        loss = step(torch.from_numpy(batch["image0"]).pin_memory(), torch.from_numpy(batch["image1"]).pin_memory(),
                    torch.from_numpy(batch["disp"]).pin_memory())
        if itr % 10 == 0:
            print("%d %g" % (itr, float(loss)))
        if itr % SAVE_INTERVAL == 0 and itr > 0:
            torch.save(model.state_dict(), os.path.join(out_dir, "model%d" % itr))
        t_iter.append(time.time() - t0)
        if itr % 100 == 1:
            print("average time per iteration: %.4fs" % (sum(t_iter) / len(t_iter)))
            t_iter = []
    torch.save(model.state_dict(), os.path.join(out_dir, "model"))


if __name__ == "__main__":
    main()
