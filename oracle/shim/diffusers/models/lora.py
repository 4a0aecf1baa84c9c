def adjust_lora_scale_text_encoder(text_encoder, lora_scale=1.0):
    return None
