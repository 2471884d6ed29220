package org.deeplearning4j.nn.conf;
public enum WorkspaceMode { NONE, ENABLED }
