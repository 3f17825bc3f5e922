python -m pytest tests/test_encoder_integration.py -m gpu -x -q -k "long_lived or batch_driver" 2>&1 | tail -2
