"""graphlearn/examples/tf"""
