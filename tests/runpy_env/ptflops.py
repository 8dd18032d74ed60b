"""test stand-in for ptflops: like the real get_model_complexity_info (trainers/trainer.py:85-90) it hooks every module and
runs ONE full forward built by `input_constructor`, so a drop-in network must survive hook injection on the DataParallel-
wrapped module.  MACs are not modelled (0); the parameter count is real."""
import torch


def get_model_complexity_info(model, input_res, print_per_layer_stat=True, as_strings=True, input_constructor=None,
                              ost=None, verbose=False, ignore_modules=(), custom_modules_hooks=None, **kw):
    handles, calls = [], [0]

    def hook(mod, inp, out):
        calls[0] += 1
    for m in model.modules():
        handles.append(m.register_forward_hook(hook))
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            if input_constructor is not None:
                model(**input_constructor(input_res))
            else:
                p = next(model.parameters())
                model(torch.ones((1,) + tuple(input_res), dtype=p.dtype, device=p.device))
    finally:
        for h in handles:
            h.remove()
        model.train(was_training)
    params = sum(p.numel() for p in model.parameters())
    return (0.0, float(params))
