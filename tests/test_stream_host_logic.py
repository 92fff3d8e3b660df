"""Host logic of the 16-bit latent / gradient streams (ops.placeholder, ops.GradChannel): no kernels, no GPU."""
import pytest
import torch


def test_placeholder_is_recognised_and_costs_no_memory():
    from gen_fvgn_steady_b200 import ops
    like = torch.zeros(3)
    p = ops.placeholder(1000, like)
    assert p.shape == (1000, 128) and p.dtype == torch.float32 and p.stride() == (0, 0)
    assert p.untyped_storage().nbytes() == 4          # one element, however many rows
    assert torch.isnan(p).all()                        # reading it is loudly wrong
    assert ops.is_placeholder(p)
    assert not ops.is_placeholder(torch.zeros(1000, 128)) and not ops.is_placeholder(None)
    assert not ops.is_placeholder(p.contiguous())      # a materialised copy is an ordinary tensor


def test_grad_channel_protocol():
    """The consumer's backward put()s the 16-bit rows of a placeholder latent, the producer's backward take()s them exactly
    once; a channel only serves the very tensors it was made for."""
    from gen_fvgn_steady_b200 import ops
    ch = ops.GradChannel()
    x, e, other = ops.placeholder(10, torch.zeros(1)), ops.placeholder(20, torch.zeros(1)), ops.placeholder(10, torch.zeros(1))
    ch.x_ref, ch.e_ref = x, e
    assert ch.serves("x", x) and ch.serves("e", e)
    assert not ch.serves("x", other) and not ch.serves("x", e) and not ch.serves("e", None)
    rows = torch.ones(10, 128, dtype=torch.float16)
    ch.put("x", rows)
    assert ch.take("x") is rows
    with pytest.raises(RuntimeError):
        ch.take("x")                                   # taken once; a second backward must not see stale rows
    with pytest.raises(RuntimeError):
        ch.take("e")                                   # never put


def test_placeholder_gradient_flows_through_autograd():
    """A placeholder carries the autograd edge: a Function that returns one receives, in its backward, whatever gradient its
    consumer returns for it (here a placeholder again: the real rows would travel through the channel)."""
    from gen_fvgn_steady_b200 import ops
    seen = {}

    class Producer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w):
            return ops.placeholder(6, w)

        @staticmethod
        def backward(ctx, g):
            seen["producer_got_placeholder"] = ops.is_placeholder(g)
            return torch.ones(1)

    class Consumer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            seen["consumer_input_is_placeholder"] = ops.is_placeholder(x)
            return torch.zeros(2)

        @staticmethod
        def backward(ctx, g):
            return ops.placeholder(6, g)

    w = torch.zeros(1, requires_grad=True)
    Consumer.apply(Producer.apply(w)).sum().backward()
    assert seen == {"consumer_input_is_placeholder": True, "producer_got_placeholder": True}
    assert float(w.grad) == 1.0


def test_grad_channel_closes_no_reference_cycle():
    """ctx -> channel -> placeholder -> grad_fn (= ctx) would be a cycle only the garbage collector frees (it slowed the
    small-mesh eager step down by 2x before the channel held its tensors weakly): with the collector off, dropping the last
    reference to the producer's output must free it at once."""
    import gc
    import weakref
    from gen_fvgn_steady_b200 import ops

    class Producer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w, ch):
            ctx.chan_out = ch
            out = ops.placeholder(6, w)
            ch.x_ref = out
            return out

        @staticmethod
        def backward(ctx, g):
            return torch.ones(1), None

    gc.collect()
    gc.disable()
    try:
        w = torch.zeros(1, requires_grad=True)
        ch = ops.GradChannel()
        out = Producer.apply(w, ch)
        assert ch.serves("x", out)
        probe = weakref.ref(out)
        del out, ch
        assert probe() is None
    finally:
        gc.enable()
