"""`regularizations` of the reference (python/regularizations.py).  The drivers index the result (`reg_loss[0]`,
python/shape_opt.py:98), so it comes back as a one-element tensor."""
import sdfrt


def eval_discrete_laplacian_reg(data, sdf=None):
    return sdfrt.regularizations.eval_discrete_laplacian_reg(data, sdf).reshape(1)
