"""Model-level drop-in surface: the colour / depth / colour+depth prediction model.

Mirrors dyn_mult_view/multi_view_model/main_model.py:12-154 (``Base_Prediction_Model``; BASELINE config 4,
conf ``tensorflowdata/cars_colordepth/conf.py``): one 5-conv pre-encoder per input modality
(``image_preprocessing`` :57-65, scopes ``pre_image0`` / ``pre_dimage0``), channel concat, shared trunk
(:105-121), ``d3_0`` widened to ``64 * num_decode`` and split (:123-137, ``split_list.pop()`` hands the
LAST group to the colour head), one decoder per output (``decode`` :68-81) with a tanh head, and
loss = L(colour) + depth_lr_factor * L(depth) (:144-154).

Feature switches are key-presence tests on ``conf`` as in the reference (``'use_color' in conf``).
Extensions (optional conf keys): image_size (128), viewpoint_dim (2), loss ('l2' reference | 'l1' per-channel L1,
tf_utils.py:22), head ('tanh' reference | 'flow': each head emits a 2-channel flow and the output is the bilinear
warp of its source modality -- the north-star's "warped RGB/depth"; precedent: multiobject_appflow.py:191-209),
grid_order, algo, seed.  The reference ignores conf['batch_size'] (main_model.py:19 hard-codes 64); here conf wins.
"""
import torch

from . import functional as F
from .model_base import ModelBase
from .optimizer import TFAdam
from .tf_utils import conv2d_msra, deconv2d_msra, flow_resample_layer, linear_msra
from .variables import VariableStore, use_store


class Base_Prediction_Model(ModelBase):
    INPUT_KEYS = ("image0", "depth0", "image1", "depth1", "disp")      # train_step(image0, dimage0, image1, dimage1, disp)

    def __init__(self, conf, load_tfrec=True, build_loss=True, device=None):
        self.conf = conf
        self.batch_size = int(conf.get("batch_size", 64))
        H = int(conf.get("image_size", 128))
        if H % 32:
            raise ValueError("image_size must be a multiple of 32")
        self.image_shape = [H, H, 3]
        self.viewpoint_dim = int(conf.get("viewpoint_dim", 2))
        self.loss_mode = conf.get("loss", "l2")
        self.head = conf.get("head", "tanh")
        self.grid_order = conf.get("grid_order", "ref_yx")
        self.algo = conf.get("algo", None)
        self.use_color = "use_color" in conf
        self.use_depth = "use_depth" in conf
        if not (self.use_color or self.use_depth):
            raise ValueError("conf needs 'use_color' and/or 'use_depth'")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.store = VariableStore(self.device, seed=int(conf.get("seed", 0)))
        self.world_size = 1
        self.gen_image1 = self.gen_dimage1 = self.loss = None
        self.optimizer = None
        B = self.batch_size
        z = lambda c: torch.zeros((B, H, H, c), dtype=torch.float32, device=self.device)
        zd = torch.zeros((B, self.viewpoint_dim), dtype=torch.float32, device=self.device)
        if self.device.type == "meta":
            with torch.no_grad(), F.meta_mode():
                self.forward(z(3), z(1), zd)
        else:
            with torch.no_grad():
                self.forward(z(3), z(1), zd)
        self.store.finalize()
        self.t_vars = self.store.trainable_vars()
        if build_loss and self.device.type != "meta":
            self.optimizer = TFAdam(self.store, conf["learning_rate"])

    # -- graph ---------------------------------------------------------------------------------
    def image_preprocessing(self, x, scope):
        """main_model.py:57-65"""
        g = self.algo
        with self.store.scope(scope):
            e0 = conv2d_msra(x, 32, 5, 5, 2, 2, "e0", act="lrelu", algo=g)
            e0_0 = conv2d_msra(e0, 32, 5, 5, 1, 1, "e0_0", act="lrelu", algo=g)
            e1 = conv2d_msra(e0_0, 32, 5, 5, 2, 2, "e1", act="lrelu", algo=g)
            e1_0 = conv2d_msra(e1, 32, 5, 5, 1, 1, "e1_0", act="lrelu", algo=g)
            return conv2d_msra(e1_0, 64, 5, 5, 2, 2, "e2", act="lrelu", algo=g)

    def decode(self, x, scope, num_channels, src=None):
        """main_model.py:68-81 (tanh head) or, with head='flow', a 2-channel flow head + warp of ``src``."""
        g = self.algo
        B, H = x.shape[0], self.image_shape[0]
        h5 = H // 32
        with self.store.scope(scope):
            d2 = deconv2d_msra(x, [B, 8 * h5, 8 * h5, 32], 5, 5, 2, 2, "d2", act="lrelu", algo=g)
            d2_0 = conv2d_msra(d2, 64, 5, 5, 1, 1, "d2_0", act="lrelu", algo=g)
            d1 = deconv2d_msra(d2_0, [B, 16 * h5, 16 * h5, 32], 5, 5, 2, 2, "d1", act="lrelu", algo=g)
            d1_0 = conv2d_msra(d1, 32, 5, 5, 1, 1, "d1_0", act="lrelu", algo=g)
            if self.head == "flow":
                flow = deconv2d_msra(d1_0, [B, H, H, 2], 5, 5, 2, 2, "d0", act=None, algo=g, out_dtype=torch.float32)
                return flow_resample_layer(src, flow, self.grid_order), flow
            gen = deconv2d_msra(d1_0, [B, H, H, num_channels], 5, 5, 2, 2, "d0", act="tanh", algo=g, out_dtype=torch.float32)
            return gen, None

    def buildModel(self, image0, dimage0, disp):
        """main_model.py:83-142"""
        g = self.algo
        B, H = image0.shape[0], image0.shape[1]
        h5 = H // 32
        concat_list = []
        if self.use_color:
            concat_list.append(self.image_preprocessing(image0, "pre_image0"))
        if self.use_depth:
            concat_list.append(self.image_preprocessing(dimage0, "pre_dimage0"))
        comb_enc = torch.cat(concat_list, dim=3) if len(concat_list) > 1 else concat_list[0]
        a = "lrelu"
        e2_0 = conv2d_msra(comb_enc, 64, 5, 5, 1, 1, "e2_0", act=a, algo=g)
        e3 = conv2d_msra(e2_0, 128, 3, 3, 2, 2, "e3", act=a, algo=g)
        e3_0 = conv2d_msra(e3, 128, 3, 3, 1, 1, "e3_0", act=a, algo=g)
        e4 = conv2d_msra(e3_0, 256, 3, 3, 2, 2, "e4", act=a, algo=g)
        e4_0 = conv2d_msra(e4, 256, 3, 3, 1, 1, "e4_0", act=a, algo=g)
        e5 = linear_msra(F.reshape(e4_0, (B, h5 * h5 * 256)), 4096, "fc1", act=a, algo=g)
        a0 = linear_msra(disp, 64, "a0", act=a, algo=g)
        a1 = linear_msra(a0, 64, "a1", act=a, algo=g)
        a2 = linear_msra(a1, 64, "a2", act=a, algo=g)
        a3 = linear_msra(torch.cat([e5, a2], dim=1), 4096, "a3", act=a, algo=g)
        a4 = linear_msra(a3, 4096, "a4", act=a, algo=g)
        a5 = linear_msra(a4, h5 * h5 * 256, "a5", act=a, algo=g)
        a5r = F.reshape(a5, (B, h5, h5, 256))
        d4 = deconv2d_msra(a5r, [B, 2 * h5, 2 * h5, 128], 3, 3, 2, 2, "d4", act=a, algo=g)
        d4_0 = conv2d_msra(d4, 128, 3, 3, 1, 1, "d4_0", act=a, algo=g)
        d3 = deconv2d_msra(d4_0, [B, 4 * h5, 4 * h5, 64], 3, 3, 2, 2, "d3", act=a, algo=g)
        num_decode = int(self.use_color) + int(self.use_depth)
        d3_0 = conv2d_msra(d3, 64 * num_decode, 5, 5, 1, 1, "d3_0", act=a, algo=g)
        split_list = list(torch.chunk(d3_0, num_decode, dim=3))
        self.flow_image1 = self.flow_dimage1 = None
        if self.use_color:       # split_list.pop(): the colour head takes the LAST channel group (main_model.py:131-133)
            self.gen_image1, self.flow_image1 = self.decode(split_list.pop(), "dec_image1", 3, image0)
        if self.use_depth:
            self.gen_dimage1, self.flow_dimage1 = self.decode(split_list.pop(), "dec_dimage1", 1, dimage0)
        assert split_list == []

    def forward(self, image0, dimage0, disp):
        """image0 [B,H,H,3], dimage0 [B,H,H,1] fp32 in [0,1], disp [B,V] -> dict(gen_image1, gen_dimage1)."""
        self.gen_image1 = self.gen_dimage1 = self.loss = None
        self.store.new_anchor()
        self.image0, self.dimage0, self.disp = image0, dimage0, disp
        with use_store(self.store):
            self.buildModel(image0, dimage0, F.to_bf16(disp))
        out = {}
        if self.use_color:
            out["gen_image1"] = self.gen_image1
        if self.use_depth:
            out["gen_dimage1"] = self.gen_dimage1
        return out

    def build_loss(self, image1, dimage1):
        """main_model.py:144-154: loss = L(colour) + L(depth) * depth_lr_factor; mean over the global batch."""
        n = image1.shape[0] * image1.shape[1] * image1.shape[2] * self.world_size
        loss = None
        if self.use_color:
            loss = F.reconstruction_loss(self.gen_image1, image1, self.loss_mode, inv_count=1.0 / n, unit_upstream=True)
        if self.use_depth:
            f = float(self.conf["depth_lr_factor"])
            ld = F.reconstruction_loss(self.gen_dimage1, dimage1, self.loss_mode, weights=[f], inv_count=1.0 / n, unit_upstream=True)
            loss = ld if loss is None else loss + ld
        self.image1, self.dimage1, self.loss = image1, dimage1, loss
        return loss

    def input_spec(self):
        B, H = self.batch_size, self.image_shape[0]
        return {"image0": (B, H, H, 3), "depth0": (B, H, H, 1), "image1": (B, H, H, 3), "depth1": (B, H, H, 1),
                "disp": (B, self.viewpoint_dim)}

    def step_loss(self, batch):
        self.forward(batch["image0"], batch["depth0"], batch["disp"])
        return self.build_loss(batch["image1"], batch["depth1"])
