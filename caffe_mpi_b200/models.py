"""Prototxt text generators for the BASELINE nets (the reference's models/*.prototxt are not available on the GPU box,
so bench.py builds the same architectures from here; tests/test_prototxt.py checks that the generated net has exactly
the inventory of the reference file when the reference tree is mounted).  Layer and blob names follow
models/resnet50/train_val.prototxt (res<stage>.<block>.conv<k>, .../bn, .skipConv, .sum, .relu, pool1/pool2, fc, loss)."""


def _conv(name, bottom, top, num_output, k, stride=1, pad=0):
    extra = (f" stride: {stride}" if stride != 1 else "") + (f" pad: {pad}" if pad else "")
    return (f'layer {{ name: "{name}" type: "Convolution" bottom: "{bottom}" top: "{top}"\n'
            f'  convolution_param {{ num_output: {num_output} kernel_size: {k}{extra} weight_filler {{ type: "msra" }} bias_term: false }} }}\n')


def _bn(name, bottom, top):
    return (f'layer {{ name: "{name}" type: "BatchNorm" bottom: "{bottom}" top: "{top}"\n'
            f'  batch_norm_param {{ moving_average_fraction: 0.9 eps: 0.0001 scale_bias: true }} }}\n')


def _relu(name, blob):
    return f'layer {{ name: "{name}" type: "ReLU" bottom: "{blob}" top: "{blob}" }}\n'


def resnet50_prototxt(batch=32, crop=224, num_classes=1000):
    s = 'name: "Resnet50"\n'
    s += (f'layer {{ name: "data" type: "Data" top: "data" top: "label" data_param {{ source: "synthetic" batch_size: {batch} backend: LMDB }}\n'
          f'  transform_param {{ crop_size: {crop} mirror: true }} include: {{ phase: TRAIN }} }}\n')
    s += _conv("conv1", "data", "conv1", 64, 7, 2, 3) + _bn("conv1/bn", "conv1", "conv1/bn") + _relu("conv1/relu", "conv1/bn")
    s += 'layer { name: "pool1" type: "Pooling" bottom: "conv1/bn" top: "pool1" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }\n'
    prev = "pool1"
    for stage, (blocks, mid, out) in zip((2, 3, 4, 5), ((3, 64, 256), (4, 128, 512), (6, 256, 1024), (3, 512, 2048))):
        for blk in range(1, blocks + 1):
            p = f"res{stage}.{blk}"
            stride = 2 if (blk == 1 and stage > 2) else 1          # stride sits on the 1x1 convs in this variant
            s += _conv(f"{p}.conv1", prev, f"{p}.conv1", mid, 1, stride) + _bn(f"{p}.conv1/bn", f"{p}.conv1", f"{p}.conv1/bn") + _relu(f"{p}.conv1/relu", f"{p}.conv1/bn")
            s += _conv(f"{p}.conv2", f"{p}.conv1/bn", f"{p}.conv2", mid, 3, 1, 1) + _bn(f"{p}.conv2/bn", f"{p}.conv2", f"{p}.conv2/bn") + _relu(f"{p}.conv2/relu", f"{p}.conv2/bn")
            s += _conv(f"{p}.conv3", f"{p}.conv2/bn", f"{p}.conv3", out, 1) + _bn(f"{p}.conv3/bn", f"{p}.conv3", f"{p}.conv3/bn")
            if blk == 1:
                s += _conv(f"{p}.skipConv", prev, f"{p}.skipConv", out, 1, stride) + _bn(f"{p}.skipConv/bn", f"{p}.skipConv", f"{p}.skipConv/bn")
                short = f"{p}.skipConv/bn"
            else:
                short = prev
            s += (f'layer {{ name: "{p}.sum" type: "Eltwise" bottom: "{p}.conv3/bn" bottom: "{short}" top: "{p}.sum" eltwise_param {{ operation: SUM }} }}\n')
            s += _relu(f"{p}.relu", f"{p}.sum")
            prev = f"{p}.sum"
    s += f'layer {{ name: "pool2" type: "Pooling" bottom: "{prev}" top: "pool2" pooling_param {{ pool: AVE kernel_size: 7 }} }}\n'
    s += (f'layer {{ name: "fc" type: "InnerProduct" bottom: "pool2" top: "fc" inner_product_param {{ num_output: {num_classes}\n'
          f'  weight_filler {{ type: "msra" }} bias_filler {{ type: "constant" value: 0 }} }} }}\n')
    s += 'layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc" bottom: "label" top: "loss" }\n'
    return s


RESNET50_SOLVER = ('base_lr: 0.001 lr_policy: "poly" power: 2.0 momentum: 0.9 weight_decay: 1e-4 max_iter: 2400000 '
                   'solver_mode: GPU random_seed: 1')    # models/resnet50/solver.prototxt (training hyper-parameters)
