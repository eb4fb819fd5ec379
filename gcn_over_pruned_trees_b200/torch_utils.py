"""Optimiser / checkpoint-config helpers the trainer needs (counterparts of /root/reference/utils/torch_utils.py:93-110,156-164)."""
import torch


def get_optimizer(name, parameters, lr, l2=0, capturable=False):
    """Same names and hyper-parameters as the reference's get_optimizer (torch_utils.py:93-106).  ``capturable``
    (the trainer passes opt['cuda']): Adam / Adamax / Adadelta keep their step counters on the device, so that
    ``optimizer.step()`` can be recorded into a CUDA graph (engine.GraphedTrainStep); same update rule."""
    cap = {'capturable': True} if capturable else {}
    if name == 'sgd':
        return torch.optim.SGD(parameters, lr=lr, weight_decay=l2)
    if name in ('adagrad', 'myadagrad'):
        # the reference's MyAdagrad = Adagrad with accumulator initialised to 0.1 (torch_utils.py:10-90)
        return torch.optim.Adagrad(parameters, lr=lr, initial_accumulator_value=0.1, eps=1e-10, weight_decay=l2)
    if name == 'adam':
        return torch.optim.Adam(parameters, weight_decay=l2, **cap)
    if name == 'adamax':
        return torch.optim.Adamax(parameters, weight_decay=l2, **cap)
    if name == 'adadelta':
        return torch.optim.Adadelta(parameters, lr=lr, weight_decay=l2, **cap)
    raise Exception("Unsupported optimizer: {}".format(name))


def change_lr(optimizer, new_lr):
    for group in optimizer.param_groups:
        group['lr'] = new_lr


def keep_partial_grad(grad, topk):
    """Zero the gradient rows >= topk (finetune only the top-N embeddings, gcn.py:83-86)."""
    assert topk < grad.size(0)
    grad.data[topk:].zero_()
    return grad


def load_config(filename):
    dump = torch.load(filename, map_location=None if torch.cuda.is_available() else torch.device('cpu'))
    return dump['config']
