"""Network plugin boundary -- host-side mirror of src/monopsr/builders/net_builder.py:9-27:

    extract_features(model, net_type, model_config, input_dict, is_training)
        -> {FEATURES_FOR_MAP: (N,48,48,128), FEATURES_FOR_BOX_3D: (N,6,6,512)}
    input_dict = {NET_IN_RGB_CROP: (N,48,48,3), NET_IN_FULL_IMG: (1,160,608,3)}

Only net_type 'resnet101_4x_squash' exists (as in the reference); `model` carries what the reference function reads from
its model object: the engine (`model.engine`, or the Engine itself) and the normalised 2-D boxes
(`model.boxes_2d_norm`, the value fed to `pl_boxes_2d_norm`).  The two ResNet-101 towers, crop-and-resize + pool,
squash and the map decoder run as the same kernels as a full forward pass (Engine.forward(features_only=True))."""
NET_IN_RGB_CROP = "net_in_rgb_crop"
NET_IN_FULL_IMG = "net_in_full_img"
FEATURES_FOR_MAP = "features_for_map"
FEATURES_FOR_BOX_3D = "features_for_box_3d"


def get_net_config(model_config):
    return getattr(model_config.net_config, model_config.net_type)


def extract_features(model, net_type, model_config, input_dict, is_training):
    if net_type != "resnet101_4x_squash":
        raise ValueError("Invalid net_type", net_type)
    engine = getattr(model, "engine", model)
    boxes = getattr(model, "boxes_2d_norm", None)
    inputs = {"rgb_crops": input_dict[NET_IN_RGB_CROP], "full_img": input_dict[NET_IN_FULL_IMG]}
    if boxes is not None:
        inputs["boxes_2d_norm"] = boxes
    engine.set_inputs(inputs)
    f = engine.forward(train=bool(is_training), features_only=True)
    return {FEATURES_FOR_MAP: f["features_for_map"], FEATURES_FOR_BOX_3D: f["features_for_box_3d"]}
