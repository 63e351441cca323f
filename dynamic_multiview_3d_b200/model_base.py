"""What every model class of the path shares: the step interface the driver and the CUDA-graph step use, and the
checkpoint surface.

The reference's driver touches a model only through ``Model(conf, load_tfrec, build_loss)``, ``model.loss``,
``model.train_op`` and the savers (dyn_mult_view/multi_view_model/train.py:57-65,117-136).  There is no TF session
here, so the same contract is:

  INPUT_KEYS            names of the tensors one ``sess.run`` consumes (the TFRecord feature names where they exist)
  input_spec()          name -> shape of one batch (float32 NHWC, values in [0,1]; viewpoint rows float32)
  train_step(batch)     one ``sess.run([loss, train_op])``: forward, backward, exchange (data parallel), Adam
  eval_loss(batch)      one ``sess.run([loss], {train_cond: 0})``: the validation pass (train.py:128-132), no update
  state_dict()/load_state_dict()   tf.train.Saver over all global variables incl. the Adam slots (train.py:70-71)

``train_step`` also accepts the class's positional form (``train_step(image0, image1, disp)`` etc.).
"""
import torch


class ModelBase(object):
    INPUT_KEYS = ()

    # -- step interface ------------------------------------------------------------------------------------
    def input_spec(self):
        raise NotImplementedError

    def _as_batch(self, args):
        if len(args) == 1 and isinstance(args[0], dict):
            return args[0]
        if len(args) != len(self.INPUT_KEYS):
            raise TypeError("%s.train_step takes a batch dict or %d tensors %s" % (type(self).__name__, len(self.INPUT_KEYS), self.INPUT_KEYS))
        return dict(zip(self.INPUT_KEYS, args))

    def step_loss(self, batch):
        """forward + loss on the training path; returns the scalar loss tensor (with its tape)."""
        raise NotImplementedError

    def train_step(self, *args):
        """One sess.run([loss, train_op]) of train.py:122: forward, backward, (gradient exchange,) Adam."""
        dp = getattr(self, "_dp", None)
        if dp is not None and hasattr(dp, "apply_deferred"):
            dp.apply_deferred()    # the late FC matrices take the previous step's update next to this step's encoder forward
        pp = getattr(self.store, "prepack", None)
        if pp is not None:
            pp.begin()             # every layer's packed weights for this step, on a side stream (functional.Prepack)
        try:
            return self._train_step_body(args)
        finally:
            if pp is not None:
                pp.end()

    def _train_step_body(self, args):
        loss = self.step_loss(self._as_batch(args))
        opt = self.optimizer
        opt.begin_step()           # single process: the big FC matrices are updated inside backward (optimizer.py)
        try:
            loss.backward()
        finally:
            opt.end_backward()
        if getattr(self, "_dp", None) is not None:
            self._dp.finish()
        opt.step()
        return loss.detach()

    def flush_updates(self):
        """Make every parameter reflect the completed steps (a deferred optimizer update may still be pending after
        train_step returns; forward passes and state_dict() do this themselves)."""
        dp = getattr(self, "_dp", None)
        if dp is not None and hasattr(dp, "flush"):
            dp.flush()

    def eval_loss(self, *args):
        """The validation pass of train.py:128-132 (val_summ_op): loss only, no gradient, no update."""
        with torch.no_grad():
            return self.step_loss(self._as_batch(args)).detach()

    def visualize(self, *a, **k):
        raise NotImplementedError("%s has no visualize(); the reference's own is broken for this class (SURVEY 3.3)" % type(self).__name__)

    # -- checkpoint surface (tf.train.Saver over GLOBAL_VARIABLES, train.py:70-71) --------------------------------
    def state_dict(self):
        """Every variable under its TF name, plus Adam's slots ``<name>/Adam`` (m), ``<name>/Adam_1`` (v) and the
        step state (beta powers, t).  Under sharded data parallelism this is a COLLECTIVE call: the fp32 masters and
        the moments live with their owning rank and are gathered first, so every rank must call it."""
        dp = getattr(self, "_dp", None)
        if dp is not None and hasattr(dp, "gather_full_state"):
            dp.gather_full_state()
        sd = self.store.state_dict()
        if self.optimizer is not None:
            for k, v in self.store.vars.items():
                if v.trainable:
                    sd[k + "/Adam"] = v.m.detach().cpu().clone()
                    sd[k + "/Adam_1"] = v.v.detach().cpu().clone()
            sd["__adam_state__"] = self.optimizer.state.detach().cpu().clone()
        return sd

    def load_state_dict(self, sd):
        dp = getattr(self, "_dp", None)
        if dp is not None and hasattr(dp, "reset_deferred"):
            dp.flush()
        self.store.load_state_dict({k: v for k, v in sd.items() if k in self.store.vars})
        if self.optimizer is not None:
            for k, v in self.store.vars.items():
                if k + "/Adam" in sd:
                    v.m.copy_(torch.as_tensor(sd[k + "/Adam"]).to(self.device))
                    v.v.copy_(torch.as_tensor(sd[k + "/Adam_1"]).to(self.device))
            if "__adam_state__" in sd:
                self.optimizer.state.copy_(torch.as_tensor(sd["__adam_state__"]).to(self.device))
