"""Whole-net CPU oracle for the trainer tests: a small layer-spec list is rendered to prototxt text (what the product
parses) and interpreted in numpy with the oracle's per-layer restatements (oracle.conv_* and oracle.layers_oracle),
including Net::Backward's diff accumulation where a blob fans out and the SGD update of Solver::Step.
TEST INFRASTRUCTURE ONLY."""
import numpy as np
import oracle
from oracle import layers_oracle as lo

POOL = {"MAX": 0, "AVE": 1}          # PoolingParameter.PoolMethod (caffe.proto)


def bottleneck(spec, name, bottom, mid, out, stride, skip):
    spec += [dict(t="conv", n=f"{name}.conv1", b=bottom, o=mid, k=1, s=stride, p=0), dict(t="bn", n=f"{name}.conv1/bn", b=f"{name}.conv1"),
             dict(t="relu", n=f"{name}.conv1/relu", b=f"{name}.conv1/bn"),
             dict(t="conv", n=f"{name}.conv2", b=f"{name}.conv1/bn", o=mid, k=3, s=1, p=1), dict(t="bn", n=f"{name}.conv2/bn", b=f"{name}.conv2"),
             dict(t="relu", n=f"{name}.conv2/relu", b=f"{name}.conv2/bn"),
             dict(t="conv", n=f"{name}.conv3", b=f"{name}.conv2/bn", o=out, k=1, s=1, p=0), dict(t="bn", n=f"{name}.conv3/bn", b=f"{name}.conv3")]
    short = bottom
    if skip:
        spec += [dict(t="conv", n=f"{name}.skipConv", b=bottom, o=out, k=1, s=stride, p=0), dict(t="bn", n=f"{name}.skipConv/bn", b=f"{name}.skipConv")]
        short = f"{name}.skipConv/bn"
    spec += [dict(t="sum", n=f"{name}.sum", b=[f"{name}.conv3/bn", short]), dict(t="relu", n=f"{name}.relu", b=f"{name}.sum")]
    return f"{name}.sum"


def mini_resnet(batch=4, size=16, classes=10, conv_bias=False):
    """conv-bn-relu-maxpool stem, a strided bottleneck with a projection shortcut, an identity bottleneck, ave-pool, fc, loss."""
    spec = [dict(t="data", n="data", shape=(batch, 3, size, size)),
            dict(t="conv", n="conv1", b="data", o=8, k=3, s=1, p=1, bias=conv_bias), dict(t="bn", n="conv1/bn", b="conv1"),
            dict(t="relu", n="conv1/relu", b="conv1/bn"), dict(t="pool", n="pool1", b="conv1/bn", m="MAX", k=3, s=2, p=0)]
    top = bottleneck(spec, "resA.1", "pool1", 4, 16, 2, True)
    top = bottleneck(spec, "resA.2", top, 4, 16, 1, False)
    hw = lo.pooled_extent(size, 3, 2, 0) // 2
    spec += [dict(t="pool", n="pool2", b=top, m="AVE", k=hw, s=1, p=0), dict(t="fc", n="fc", b="pool2", o=classes),
             dict(t="loss", n="loss", b=["fc", "label"])]
    return spec


def lenet(batch=64, classes=10):
    """examples/mnist/lenet_train_test.prototxt (BASELINE configs[0]): conv(20,5) pool conv(50,5) pool ip(500) relu ip(10) loss."""
    return [dict(t="data", n="data", shape=(batch, 1, 28, 28)),
            dict(t="conv", n="conv1", b="data", o=20, k=5, s=1, p=0, bias=True), dict(t="pool", n="pool1", b="conv1", m="MAX", k=2, s=2, p=0),
            dict(t="conv", n="conv2", b="pool1", o=50, k=5, s=1, p=0, bias=True), dict(t="pool", n="pool2", b="conv2", m="MAX", k=2, s=2, p=0),
            dict(t="fc", n="ip1", b="pool2", o=500), dict(t="relu", n="relu1", b="ip1"), dict(t="fc", n="ip2", b="ip1", o=classes),
            dict(t="loss", n="loss", b=["ip2", "label"])]


def mini_inception(batch=4, size=12, classes=10):
    """The layer kinds AlexNet / GoogLeNet / VGG-16 add to the ResNet set, in one small net: grouped conv (AlexNet conv2), LRN
    across channels, an inception module (1x1 / 3x3 / 5x5 / pool-proj branches joined by Concat, every branch conv with bias +
    in-place ReLU), an auxiliary classifier with loss_weight 0.3 and Dropout, the main classifier behind Dropout."""
    spec = [dict(t="data", n="data", shape=(batch, 4, size, size)),
            dict(t="conv", n="conv1", b="data", o=8, k=3, s=1, p=1, g=2, bias=True), dict(t="relu", n="relu1", b="conv1"),
            dict(t="lrn", n="norm1", b="conv1", size=5, alpha=1e-2, beta=0.75, k=1.0),
            dict(t="pool", n="pool1", b="norm1", m="MAX", k=3, s=2, p=0)]
    src = "pool1"
    for br, (k, o, p) in (("1x1", (1, 6, 0)), ("3x3", (3, 8, 1)), ("5x5", (5, 4, 2))):
        spec += [dict(t="conv", n=f"inc/{br}", b=src, o=o, k=k, s=1, p=p, bias=True), dict(t="relu", n=f"inc/relu_{br}", b=f"inc/{br}")]
    spec += [dict(t="pool", n="inc/pool", b=src, m="MAX", k=3, s=1, p=1),
             dict(t="conv", n="inc/pool_proj", b="inc/pool", o=6, k=1, s=1, p=0, bias=True), dict(t="relu", n="inc/relu_pp", b="inc/pool_proj"),
             dict(t="concat", n="inc/output", b=["inc/1x1", "inc/3x3", "inc/5x5", "inc/pool_proj"]),
             # auxiliary head (GoogLeNet loss1/*): ave pool, fc, relu, dropout, classifier, loss_weight 0.3
             dict(t="pool", n="aux/pool", b="inc/output", m="AVE", k=3, s=2, p=0),
             dict(t="fc", n="aux/fc", b="aux/pool", o=16), dict(t="relu", n="aux/relu", b="aux/fc"),
             dict(t="dropout", n="aux/drop", b="aux/fc", ratio=0.5),
             dict(t="fc", n="aux/cls", b="aux/fc", o=classes), dict(t="loss", n="aux/loss", b=["aux/cls", "label"], w=0.3, top="aux/loss1"),
             # main head
             dict(t="pool", n="pool5", b="inc/output", m="AVE", k=5, s=1, p=0),
             dict(t="dropout", n="drop5", b="pool5", ratio=0.4),
             dict(t="fc", n="cls", b="pool5", o=classes), dict(t="loss", n="loss", b=["cls", "label"], w=1.0)]
    return spec


def to_prototxt(spec, eps=1e-4, maf=0.9):
    s = 'name: "mini"\n'
    for L in spec:
        t, n = L["t"], L["n"]
        if t == "data":
            N, Cc, H, W = L["shape"]
            s += (f'layer {{ name: "data" type: "Input" top: "data" top: "label" input_param {{ shape {{ dim: {N} dim: {Cc} dim: {H} dim: {W} }} '
                  f'shape {{ dim: {N} }} }} }}\n')
        elif t == "conv":
            grp = f' group: {L["g"]}' if L.get("g", 1) != 1 else ""
            s += (f'layer {{ name: "{n}" type: "Convolution" bottom: "{L["b"]}" top: "{n}" convolution_param {{ num_output: {L["o"]} '
                  f'kernel_size: {L["k"]} stride: {L["s"]} pad: {L["p"]}{grp} bias_term: {"true" if L.get("bias") else "false"} weight_filler {{ type: "msra" }} }} }}\n')
        elif t == "bn":
            s += (f'layer {{ name: "{n}" type: "BatchNorm" bottom: "{L["b"]}" top: "{n}" batch_norm_param {{ moving_average_fraction: {maf} '
                  f'eps: {eps} scale_bias: true }} }}\n')
        elif t == "relu":
            s += f'layer {{ name: "{n}" type: "ReLU" bottom: "{L["b"]}" top: "{L["b"]}" }}\n'
        elif t == "pool":
            s += (f'layer {{ name: "{n}" type: "Pooling" bottom: "{L["b"]}" top: "{n}" pooling_param {{ pool: {L["m"]} kernel_size: {L["k"]} '
                  f'stride: {L["s"]} pad: {L["p"]} }} }}\n')
        elif t == "sum":
            s += f'layer {{ name: "{n}" type: "Eltwise" bottom: "{L["b"][0]}" bottom: "{L["b"][1]}" top: "{n}" eltwise_param {{ operation: SUM }} }}\n'
        elif t == "fc":
            s += (f'layer {{ name: "{n}" type: "InnerProduct" bottom: "{L["b"]}" top: "{n}" inner_product_param {{ num_output: {L["o"]} '
                  f'weight_filler {{ type: "msra" }} bias_filler {{ type: "constant" value: 0 }} }} }}\n')
        elif t == "loss":
            lw = f' loss_weight: {L["w"]}' if "w" in L else ""
            s += f'layer {{ name: "{n}" type: "SoftmaxWithLoss" bottom: "{L["b"][0]}" bottom: "{L["b"][1]}" top: "{L.get("top", n)}"{lw} }}\n'
        elif t == "lrn":
            s += (f'layer {{ name: "{n}" type: "LRN" bottom: "{L["b"]}" top: "{n}" lrn_param {{ local_size: {L["size"]} alpha: {L["alpha"]} '
                  f'beta: {L["beta"]} k: {L["k"]} }} }}\n')
        elif t == "dropout":
            s += f'layer {{ name: "{n}" type: "Dropout" bottom: "{L["b"]}" top: "{L["b"]}" dropout_param {{ dropout_ratio: {L["ratio"]} }} }}\n'
        elif t == "concat":
            s += f'layer {{ name: "{n}" type: "Concat" ' + " ".join(f'bottom: "{b}"' for b in L["b"]) + f' top: "{n}" }}\n'
    return s


def param_shapes(spec):
    """Learnable blobs in Net::AppendParam order: conv w (,b), bn scale, bn shift, fc w, fc b."""
    shapes, chan = [], {}
    for L in spec:
        t, n = L["t"], L["n"]
        if t == "data":
            chan["data"] = L["shape"]
        elif t == "conv":
            N, Cc, H, W = chan[L["b"]]
            prm = oracle.ConvParams.make(N, Cc, H, W, L["o"], L["k"], L["s"], L["p"], 1, L.get("g", 1), bool(L.get("bias")))
            L["prm"] = prm
            shapes.append((n, "w", prm.w_shape()))
            if L.get("bias"):
                shapes.append((n, "b", (L["o"],)))
            chan[n] = prm.y_shape()
        elif t == "bn":
            chan[n] = chan[L["b"]]
            shapes += [(n, "scale", (chan[n][1],)), (n, "shift", (chan[n][1],))]
        elif t == "pool":
            N, Cc, H, W = chan[L["b"]]
            chan[n] = (N, Cc, lo.pooled_extent(H, L["k"], L["s"], L["p"]), lo.pooled_extent(W, L["k"], L["s"], L["p"]))
        elif t == "sum":
            chan[n] = chan[L["b"][0]]
        elif t == "lrn":
            chan[n] = chan[L["b"]]
        elif t == "concat":
            x0 = chan[L["b"][0]]
            chan[n] = (x0[0], sum(chan[b][1] for b in L["b"])) + tuple(x0[2:])
        elif t == "fc":
            x = chan[L["b"]]
            shapes += [(n, "w", (L["o"], int(np.prod(x[1:])))), (n, "b", (L["o"],))]
            chan[n] = (x[0], L["o"])
    return shapes


def forward_backward(spec, params, data, label, eps=1e-4, seed=1701, iteration=0):
    """params: list of arrays in param_shapes order.  Returns (loss of the LAST loss layer, grads list, blobs dict of forward
    values, diffs dict).  Dropout masks are the product's counter-based ones (layers_oracle.dropout_mask): layer li of a
    trainer seeded `seed` draws from stream seed * 0x100000001B3 + li at offset iteration * count."""
    v, saved, pi, pidx = {"data": data, "label": label}, {}, 0, {}
    for li, L in enumerate(spec):
        t, n = L["t"], L["n"]
        if t == "conv":
            w = params[pi]; pidx[n] = pi; pi += 1
            b = None
            if L.get("bias"):
                b = params[pi]; pi += 1
            saved[n] = v[L["b"]]
            v[n] = oracle.conv_forward(L["prm"], v[L["b"]], w, b)
        elif t == "bn":
            pidx[n] = pi
            y, xn, _, inv, _, _ = lo.bn_forward_train(v[L["b"]], params[pi], params[pi + 1], eps, 0.9, None, None, True)
            pi += 2
            saved[n] = (xn, inv)
            v[n] = y
        elif t == "relu":
            v[L["b"]] = lo.relu_forward(v[L["b"]])
        elif t == "pool":
            saved[n] = v[L["b"]].shape
            v[n], saved[n + "#m"] = lo.pool_forward(v[L["b"]], POOL[L["m"]], (L["k"],) * 2, (L["s"],) * 2, (L["p"],) * 2)
        elif t == "sum":
            v[n] = v[L["b"][0]] + v[L["b"][1]]
        elif t == "fc":
            pidx[n] = pi; pi += 2
            saved[n] = v[L["b"]]
            v[n] = lo.ip_forward(v[L["b"]], params[pidx[n]], params[pidx[n] + 1])
        elif t == "loss":
            saved[n], loss = lo.softmax_loss_forward(v[L["b"][0]], label)
            v[L.get("top", n)] = np.float32(loss)
        elif t == "lrn":
            v[n], saved[n] = lo.lrn_forward(v[L["b"]], L["size"], L["alpha"], L["beta"], L["k"])
            saved[n + "#x"] = v[L["b"]]
        elif t == "dropout":
            x = v[L["b"]]
            s64 = (int(seed) * 0x100000001B3 + li) & 0xFFFFFFFFFFFFFFFF
            m = lo.dropout_mask(x.size, L["ratio"], s64, iteration * x.size).reshape(x.shape)
            saved[n] = m
            v[L["b"]] = x * m
        elif t == "concat":
            v[n] = np.concatenate([v[b] for b in L["b"]], axis=1)
    grads = [np.zeros_like(p) for p in params]
    d = {}

    def acc(name, g):
        d[name] = g if name not in d else d[name] + g

    for L in reversed(spec):
        t, n = L["t"], L["n"]
        if t == "loss":
            acc(L["b"][0], lo.softmax_loss_backward(saved[n], label, L.get("w", 1.0)))
        elif t == "lrn":
            acc(L["b"], lo.lrn_backward(saved[n + "#x"], v[n], saved[n], d[n], L["size"], L["alpha"], L["beta"]))
        elif t == "dropout":
            d[L["b"]] = d[L["b"]] * saved[n]
        elif t == "concat":
            c0 = 0
            for b in L["b"]:
                cb = v[b].shape[1]
                acc(b, d[n][:, c0:c0 + cb])
                c0 += cb
        elif t == "fc":
            dw, db, dx = lo.ip_backward(saved[n], params[pidx[n]], d[n])
            grads[pidx[n]], grads[pidx[n] + 1] = dw.reshape(params[pidx[n]].shape), db
            acc(L["b"], dx)
        elif t == "sum":
            acc(L["b"][0], d[n]); acc(L["b"][1], d[n])
        elif t == "pool":
            acc(L["b"], lo.pool_backward(d[n], saved[n + "#m"], saved[n], POOL[L["m"]], (L["k"],) * 2, (L["s"],) * 2, (L["p"],) * 2))
        elif t == "relu":
            d[L["b"]] = lo.relu_backward(d[L["b"]], v[L["b"]])
        elif t == "bn":
            xn, inv = saved[n]
            dg, dbt, dx = lo.bn_backward(d[n], xn, params[pidx[n]], inv)
            grads[pidx[n]], grads[pidx[n] + 1] = dg, dbt
            acc(L["b"], dx)
        elif t == "conv":
            want_dx = L["b"] != "data"
            dw, db, dx = oracle.conv_backward(L["prm"], saved[n], params[pidx[n]], d[n], want_dx=want_dx)
            grads[pidx[n]] = dw
            if L.get("bias"):
                grads[pidx[n] + 1] = db
            if want_dx:
                acc(L["b"], dx)
    return float(loss), grads, v, d


def sgd_steps(spec, params, data, label, steps, base_lr, momentum, wd):
    """Solver::Step x steps with lr_policy fixed; returns (losses, params, history)."""
    params = [p.copy() for p in params]
    hist = [np.zeros_like(p) for p in params]
    losses = []
    for _ in range(steps):
        loss, grads, _, _ = forward_backward(spec, params, data, label)
        losses.append(loss)
        for i, g in enumerate(grads):
            _, w, h = oracle.sgd_update(g, params[i], hist[i], momentum, base_lr, wd)
            params[i], hist[i] = w.reshape(params[i].shape), h.reshape(params[i].shape)
    return losses, params, hist
