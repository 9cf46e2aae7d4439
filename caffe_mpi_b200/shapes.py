"""Convolution-layer inventories of the BASELINE.json configs, derived from the
reference's models/*.prototxt with the shape rule of conv_layer.cpp:7-22.  Data only.

Each entry: (count, C, H, O, k, s, p, G, bias)  (square maps / kernels)."""

RESNET50 = [  # models/resnet50/train_val.prototxt, all bias_term:false; 53 convs, 20 unique
    (1, 3, 224, 64, 7, 2, 3, 1, False),
    (1, 64, 56, 64, 1, 1, 0, 1, False),
    (3, 64, 56, 64, 3, 1, 1, 1, False),
    (4, 64, 56, 256, 1, 1, 0, 1, False),
    (2, 256, 56, 64, 1, 1, 0, 1, False),
    (1, 256, 56, 128, 1, 2, 0, 1, False),
    (1, 256, 56, 512, 1, 2, 0, 1, False),
    (4, 128, 28, 128, 3, 1, 1, 1, False),
    (4, 128, 28, 512, 1, 1, 0, 1, False),
    (3, 512, 28, 128, 1, 1, 0, 1, False),
    (1, 512, 28, 256, 1, 2, 0, 1, False),
    (1, 512, 28, 1024, 1, 2, 0, 1, False),
    (6, 256, 14, 256, 3, 1, 1, 1, False),
    (6, 256, 14, 1024, 1, 1, 0, 1, False),
    (5, 1024, 14, 256, 1, 1, 0, 1, False),
    (1, 1024, 14, 512, 1, 2, 0, 1, False),
    (1, 1024, 14, 2048, 1, 2, 0, 1, False),
    (3, 512, 7, 512, 3, 1, 1, 1, False),
    (3, 512, 7, 2048, 1, 1, 0, 1, False),
    (2, 2048, 7, 512, 1, 1, 0, 1, False),
]

ALEXNET = [  # models/bvlc_alexnet/train_val.prototxt (227x227), all bias
    (1, 3, 227, 96, 11, 4, 0, 1, True),
    (1, 96, 27, 256, 5, 1, 2, 2, True),
    (1, 256, 13, 384, 3, 1, 1, 1, True),
    (1, 384, 13, 384, 3, 1, 1, 2, True),
    (1, 384, 13, 256, 3, 1, 1, 2, True),
]

VGG16 = [  # models/vgg16/train_val.prototxt, all k3 s1 p1, bias
    (1, 3, 224, 64, 3, 1, 1, 1, True), (1, 64, 224, 64, 3, 1, 1, 1, True),
    (1, 64, 112, 128, 3, 1, 1, 1, True), (1, 128, 112, 128, 3, 1, 1, 1, True),
    (1, 128, 56, 256, 3, 1, 1, 1, True), (2, 256, 56, 256, 3, 1, 1, 1, True),
    (1, 256, 28, 512, 3, 1, 1, 1, True), (2, 512, 28, 512, 3, 1, 1, 1, True),
    (3, 512, 14, 512, 3, 1, 1, 1, True),
]

LENET = [  # examples/mnist/lenet_train_test.prototxt
    (1, 1, 28, 20, 5, 1, 0, 1, True),
    (1, 20, 12, 50, 5, 1, 0, 1, True),
]

# learnable-parameter element counts outside the conv layers (fc + BN scale/bias), SURVEY Appendix A
EXTRA_PARAMS = {"resnet50": 2048 * 1000 + 1000 + 53120, "alexnet": 9216 * 4096 + 4096 + 4096 * 4096 + 4096 + 4096 * 1000 + 1000,
                "vgg16": 25088 * 4096 + 4096 + 4096 * 4096 + 4096 + 4096 * 1000 + 1000, "lenet": 800 * 500 + 500 + 500 * 10 + 10}

MODELS = {"resnet50": RESNET50, "alexnet": ALEXNET, "vgg16": VGG16, "lenet": LENET}


def conv_flops_per_image(layers, train=True):
    """2*O*(C/g)*k*k*Ho*Wo per pass; training = fwd + wgrad + dgrad, no dgrad for the first (data) layer."""
    fwd, first = 0, None
    for (cnt, C, H, O, k, s, p, G, _b) in layers:
        Ho = (H + 2 * p - k) // s + 1
        f = 2 * O * (C // G) * k * k * Ho * Ho
        if first is None:
            first = f
        fwd += cnt * f
    return (3 * fwd - first) if train else fwd
